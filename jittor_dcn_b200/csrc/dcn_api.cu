// dcn_api.cu — the C ABI declared in include/dcn_b200.h: validation, workspace carving and
// dispatch to a kernel family.  No torch types, no allocation, no device synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "dcn_common.cuh"
#include "dcn_umma.h"
#include "dcn_umma_common.cuh"

namespace dcn {

static thread_local char g_err[512] = "";
// process-wide: autograd runs the backward on its own thread, bench.py reads from the main one
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return DCN_ERR_CUDA;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static int env_set(const char* name) { return getenv(name) != nullptr; }

const Knobs& knobs() {
  // C++11 magic static: initialised once, thread-safe
  static const Knobs k = [] {
    Knobs v;
    v.fwd_no_tma_out = env_set("DCN_FWD_NO_TMA_OUT");
    v.fwd_stages = env_int("DCN_FWD_STAGES", 3);
    v.fwd_no_kperm = env_set("DCN_FWD_NO_KPERM");
    v.fwd_no_split88 = env_set("DCN_FWD_NO_SPLIT88");
    v.bwd_no_resident = env_set("DCN_BWD_NO_RESIDENT");
    v.bwd_no_ring1 = env_set("DCN_BWD_NO_RING1");
    v.bwd_slice_cb = env_int("DCN_BWD_SLICE_CB", 6);
    v.bwd_no_fuse = env_int("DCN_BWD_NO_FUSE", 0);
    v.bwd_gbuf1 = env_int("DCN_BWD_GBUF", 0) == 1;
    v.bwd_data_simt = env_int("DCN_BWD_DATA_SIMT", 0);
    v.conv_off = env_set("DCN_CONV_OFF");
    v.gemm_off = env_set("DCN_GEMM_OFF");
    v.gemm_sgemm = env_set("DCN_GEMM_SGEMM");
    v.conv_small_c = env_set("DCN_CONV_SMALL_C");
    v.conv_small_off = env_set("DCN_CONV_SMALL_OFF");
    v.conv_debug = env_set("DCN_CONV_DEBUG");
    v.conv_wstream = env_set("DCN_CONV_WSTREAM");
    return v;
  }();
  return k;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- per-kernel event profiler ------------------------------------------------------
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfRec>* g_prof = nullptr;
static std::mutex g_prof_mu;
static thread_local int g_prof_slot = -1;  // record opened by this thread's KernelScope

void profile_mark(const char* name, cudaStream_t st, bool begin) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof) return;
  if (begin) {
    ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    g_prof->push_back(r);
    g_prof_slot = (int)g_prof->size() - 1;
  } else if (g_prof_slot >= 0 && g_prof_slot < (int)g_prof->size()) {
    cudaEventRecord((*g_prof)[g_prof_slot].b, st);
    g_prof_slot = -1;
  }
}

static int check_ptr(const void* p, const char* name, bool required = true) {
  if (!p) {
    if (!required) return DCN_OK;
    set_error("%s is NULL", name);
    return DCN_ERR_NULL_POINTER;
  }
  if ((uintptr_t)p & 15u) {
    set_error("%s (%p) is not 16-byte aligned", name, p);
    return DCN_ERR_MISALIGNED;
  }
  return DCN_OK;
}

static int geo_or_error(const DcnShape* s, Geo* g) {
  int rc = make_geo(s, g);
  if (rc == DCN_ERR_NULL_POINTER) set_error("DcnShape is NULL");
  if (rc == DCN_ERR_BAD_SHAPE)
    set_error("bad shape B=%d C=%d O=%d H=%d W=%d k=(%d,%d) s=(%d,%d) p=(%d,%d) variant=%d", s->B,
              s->C, s->O, s->H, s->W, s->kh, s->kw, s->sh, s->sw, s->ph, s->pw, s->variant);
  if (rc == DCN_OK && s->operand != DCN_OPERAND_FP32 && s->operand != DCN_OPERAND_BF16) {
    set_error("unknown operand mode %d", s->operand);
    rc = DCN_ERR_UNSUPPORTED;
  }
  return rc;
}

static size_t plan_bytes(const Geo& g) { return align_up(sizeof(Tap) * (size_t)g.B * g.P, 256); }

// Generic kernels in bf16 storage mode: fp32 scratch copies of the bfloat16 operands, behind the plan.
static size_t f32_copy_bytes(size_t n) { return align_up(sizeof(float) * n, 256); }
static size_t n_x(const Geo& g) { return (size_t)g.B * g.C * g.H * g.W; }
static size_t n_w(const Geo& g) { return (size_t)g.O * g.K; }
static size_t n_out(const Geo& g) { return (size_t)g.B * g.O * g.HW; }
static size_t simt_workspace(const Geo& g, int operand, int phase) {
  size_t b = plan_bytes(g);
  if (operand == DCN_OPERAND_BF16) {
    b += f32_copy_bytes(n_x(g)) + f32_copy_bytes(n_w(g));
    if (phase == DCN_PHASE_BACKWARD) b += f32_copy_bytes(n_out(g));
  }
  return b;
}

size_t bn_workspace_bytes(int C);
int bn_relu_forward_staged(const Geo& g, const Tiling& t, int training, const float* x, const float* gamma,
                           const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                           float* xt, float* saved, void* workspace, cudaStream_t st);
int bn_relu_backward_staged(const Geo& g, const Tiling& t, int training, const float* x, const float* gxt,
                            const float* saved, float* grad_x, float* grad_gamma, float* grad_beta, void* workspace,
                            cudaStream_t st);

int bn_relu_forward(int B, int C, int HW, int training, const float* x, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, float* y, float* saved,
                    void* workspace, cudaStream_t st);
int bn_relu_backward(int B, int C, int HW, int training, const float* x, const float* grad_y, const float* saved,
                     float* grad_x, float* grad_gamma, float* grad_beta, void* workspace, cudaStream_t st);

static int bn_shape_ok(int B, int C, int HW) {
  if (B <= 0 || C <= 0 || HW <= 0 || (long long)B * C * HW > (1LL << 40) || (long long)B * HW > 0x7fffffffLL) {
    set_error("bad batch-norm shape B=%d C=%d HW=%d", B, C, HW);
    return DCN_ERR_BAD_SHAPE;
  }
  return DCN_OK;
}

static bool use_umma(const DcnShape* s, const Geo& g, int phase) {
  if (s->flags & DCN_FLAG_FORCE_SIMT) return false;
  return umma_supported(g, s->operand, phase);
}

// Torch column layout with too few channels per sampling point for the fused kernels (dcn_gemm_path.cu):
// materialised samples + plain GEMMs
bool gemm_path_supported(const Geo& g, int operand);
size_t gemm_path_workspace(const Geo& g, int operand, int phase);
int gemm_path_forward(const Geo& g, int operand, const void* x, const float* off, const void* wt, const float* bias,
                      float* out, void* workspace, cudaStream_t st);
int gemm_path_backward(const Geo& g, int operand, int flags, const void* x, const float* off, const void* wt,
                       const void* gout, float* gx, float* goff, float* gw, float* gb, void* workspace, cudaStream_t st);
static bool use_gemm(const DcnShape* s, const Geo& g, int phase) {
  if ((s->flags & DCN_FLAG_FORCE_SIMT) || (phase != DCN_PHASE_FORWARD && phase != DCN_PHASE_BACKWARD)) return false;
  return !umma_supported(g, s->operand, phase) && gemm_path_supported(g, s->operand);
}

// whole-layer calls (offset conv + DCN span): fp32 operands on the tensor path only
static bool layer_ok(const DcnShape* s, const Geo& g, int phase) {
  if ((s->flags & DCN_FLAG_FORCE_SIMT) || s->operand != DCN_OPERAND_FP32) return false;
  if (!umma_supported(g, DCN_OPERAND_FP32, DCN_PHASE_FORWARD) || !umma_offset_conv_fwd_supported(g)) return false;
  if (phase == DCN_PHASE_LAYER_FORWARD) return true;
  return umma_layer_bwd_supported(g);
}

static size_t layer_workspace(const Geo& g, int operand, int phase) {
  const size_t f1 = umma_offset_conv_fwd_workspace(g), f2 = umma_workspace_bytes(g, operand, DCN_PHASE_FORWARD);
  const size_t f = f1 > f2 ? f1 : f2;
  if (phase == DCN_PHASE_LAYER_FORWARD) return f;
  return umma_layer_bwd_workspace(g);
}

}  // namespace dcn

using namespace dcn;

extern "C" {

int dcn_version(void) { return DCN_B200_VERSION; }

const char* dcn_last_error(void) { return g_err; }

const char* dcn_status_string(int status) {
  switch (status) {
    case DCN_OK: return "ok";
    case DCN_ERR_BAD_SHAPE: return "bad shape";
    case DCN_ERR_NULL_POINTER: return "null pointer";
    case DCN_ERR_MISALIGNED: return "misaligned pointer";
    case DCN_ERR_WORKSPACE: return "workspace too small";
    case DCN_ERR_CUDA: return "CUDA error";
    case DCN_ERR_UNSUPPORTED: return "unsupported configuration";
    case DCN_ERR_NCCL: return "NCCL error";
    default: return "unknown status";
  }
}

int dcn_output_hw(const DcnShape* s, int32_t* h_out, int32_t* w_out) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  if (h_out) *h_out = g.Ho;
  if (w_out) *w_out = g.Wo;
  return DCN_OK;
}

size_t dcn_workspace_bytes(const DcnShape* s, int phase) {
  Geo g;
  if (geo_or_error(s, &g)) return 0;
  if (phase == DCN_PHASE_CORNERS) return 0;
  if (phase == DCN_PHASE_LAYER_FORWARD || phase == DCN_PHASE_LAYER_BACKWARD)
    return layer_ok(s, g, phase) ? layer_workspace(g, s->operand, phase) : 0;
  if (use_umma(s, g, phase)) return umma_workspace_bytes(g, s->operand, phase);
  if (use_gemm(s, g, phase)) return gemm_path_workspace(g, s->operand, phase);
  return simt_workspace(g, s->operand, phase);
}

const char* dcn_path_name(const DcnShape* s, int phase) {
  Geo g;
  if (geo_or_error(s, &g)) return "invalid";
  if (phase == DCN_PHASE_LAYER_FORWARD || phase == DCN_PHASE_LAYER_BACKWARD)
    return layer_ok(s, g, phase) ? "umma" : "unsupported";
  return use_umma(s, g, phase) ? "umma" : (use_gemm(s, g, phase) ? "gemm" : "simt");
}

int dcn_profile_begin(void) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof) g_prof = new std::vector<ProfRec>();
  for (auto& r : *g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof->clear();
  g_prof_on = true;
  return DCN_OK;
}

int dcn_profile_end(char* out, size_t cap) {
  g_prof_on = false;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (!g_prof) return DCN_OK;
  std::map<std::string, std::pair<int, double>> agg;
  std::vector<std::string> order;
  for (auto& r : *g_prof) {
    float ms = 0.f;
    if (r.b && cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto it = agg.find(r.name);
      if (it == agg.end()) {
        order.push_back(r.name);
        agg[r.name] = {1, (double)ms};
      } else {
        it->second.first += 1;
        it->second.second += ms;
      }
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof->clear();
  size_t used = 0;
  if (out && cap) out[0] = 0;
  for (auto& n : order) {
    char line[256];
    int len = snprintf(line, sizeof line, "%s %d %.6f\n", n.c_str(), agg[n].first, agg[n].second);
    if (out && used + (size_t)len + 1 < cap) {
      memcpy(out + used, line, (size_t)len + 1);
      used += (size_t)len;
    }
  }
  return DCN_OK;
}

uint64_t dcn_launch_count(void) { return g_launches.load(); }
void dcn_launch_count_reset(void) { g_launches.store(0); }

int dcn_forward(const DcnShape* s, const void* x, const void* offset, const void* weight,
                const void* bias, void* out, void* workspace, size_t workspace_bytes,
                void* stream) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(offset, "offset")) ||
      (rc = check_ptr(weight, "weight")) || (rc = check_ptr(bias, "bias", false)) ||
      (rc = check_ptr(out, "out")))
    return rc;
  const size_t need = dcn_workspace_bytes(s, DCN_PHASE_FORWARD);
  if (need && (rc = check_ptr(workspace, "workspace"))) return rc;
  if (workspace_bytes < need) {
    set_error("forward workspace: have %zu bytes, need %zu", workspace_bytes, need);
    return DCN_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (use_umma(s, g, DCN_PHASE_FORWARD))
    return umma_forward(g, s->operand, s->flags, x, (const float*)offset, weight, (const float*)bias,
                        out, workspace, st);
  if (use_gemm(s, g, DCN_PHASE_FORWARD))
    return gemm_path_forward(g, s->operand, x, (const float*)offset, weight, (const float*)bias, (float*)out, workspace,
                             st);
  Tap* plan = (Tap*)workspace;
  if (s->operand == DCN_OPERAND_BF16) {
    // shapes the tensor path does not tile: widen the bf16 operands once, run the fp32 kernels
    float* xf = (float*)((uint8_t*)workspace + plan_bytes(g));
    float* wf = (float*)((uint8_t*)xf + f32_copy_bytes(n_x(g)));
    if ((rc = launch_widen_bf16(x, xf, n_x(g), st)) || (rc = launch_widen_bf16(weight, wf, n_w(g), st))) return rc;
    x = xf;
    weight = wf;
  }
  if ((rc = launch_plan(g, (const float*)offset, plan, st))) return rc;
  return simt_forward(g, (const float*)x, plan, (const float*)weight, (const float*)bias, (float*)out,
                      st);
}

int dcn_backward(const DcnShape* s, const void* x, const void* offset, const void* weight,
                 const void* grad_out, void* grad_x, void* grad_offset, void* grad_weight,
                 void* grad_bias, void* workspace, size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  const bool want_gx = !(s->flags & DCN_FLAG_NO_GRAD_X);
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(offset, "offset")) ||
      (rc = check_ptr(weight, "weight")) || (rc = check_ptr(grad_out, "grad_out")) ||
      (rc = check_ptr(grad_x, "grad_x", want_gx)) || (rc = check_ptr(grad_offset, "grad_offset")) ||
      (rc = check_ptr(grad_weight, "grad_weight")) || (rc = check_ptr(grad_bias, "grad_bias", false)))
    return rc;
  const size_t need = dcn_workspace_bytes(s, DCN_PHASE_BACKWARD);
  if (need && (rc = check_ptr(workspace, "workspace"))) return rc;
  if (workspace_bytes < need) {
    set_error("backward workspace: have %zu bytes, need %zu", workspace_bytes, need);
    return DCN_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (use_umma(s, g, DCN_PHASE_BACKWARD))
    return umma_backward(g, s->operand, s->flags, x, (const float*)offset, weight, grad_out,
                         (float*)grad_x, (float*)grad_offset, (float*)grad_weight,
                         (float*)grad_bias, workspace, st);
  if (use_gemm(s, g, DCN_PHASE_BACKWARD))
    return gemm_path_backward(g, s->operand, s->flags, x, (const float*)offset, weight, grad_out, (float*)grad_x,
                              (float*)grad_offset, (float*)grad_weight, (float*)grad_bias, workspace, st);
  Tap* plan = (Tap*)workspace;
  if (s->operand == DCN_OPERAND_BF16) {
    float* xf = (float*)((uint8_t*)workspace + plan_bytes(g));
    float* wf = (float*)((uint8_t*)xf + f32_copy_bytes(n_x(g)));
    float* gf = (float*)((uint8_t*)wf + f32_copy_bytes(n_w(g)));
    if ((rc = launch_widen_bf16(x, xf, n_x(g), st)) || (rc = launch_widen_bf16(weight, wf, n_w(g), st)) ||
        (rc = launch_widen_bf16(grad_out, gf, n_out(g), st)))
      return rc;
    x = xf;
    weight = wf;
    grad_out = gf;
  }
  if ((rc = launch_plan(g, (const float*)offset, plan, st))) return rc;
  return simt_backward(g, s->flags, (const float*)x, plan, (const float*)weight,
                       (const float*)grad_out, (float*)grad_x, (float*)grad_offset,
                       (float*)grad_weight, (float*)grad_bias, st);
}

int dcn_offset_conv_forward(const DcnShape* s, const void* x, const void* offset_weight, const void* offset_bias,
                            void* offset, void* workspace, size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(offset_weight, "offset_weight")) ||
      (rc = check_ptr(offset_bias, "offset_bias", false)) || (rc = check_ptr(offset, "offset")) ||
      (rc = check_ptr(workspace, "workspace")))
    return rc;
  if (!layer_ok(s, g, DCN_PHASE_LAYER_FORWARD)) {
    set_error("offset conv on the engine: shape / operand / flags not supported (dcn_path_name(s, DCN_PHASE_LAYER_FORWARD))");
    return DCN_ERR_UNSUPPORTED;
  }
  const size_t need = umma_offset_conv_fwd_workspace(g);
  if (workspace_bytes < need) {
    set_error("offset conv workspace: have %zu bytes, need %zu", workspace_bytes, need);
    return DCN_ERR_WORKSPACE;
  }
  return umma_offset_conv_forward(g, x, (const float*)offset_weight, (const float*)offset_bias, (float*)offset,
                                  workspace, (cudaStream_t)stream, !(s->flags & DCN_FLAG_XT_STAGED));
}

int dcn_layer_forward(const DcnShape* s, const void* x, const void* offset_weight, const void* offset_bias,
                      const void* weight, const void* bias, void* offset, void* out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  if ((rc = check_ptr(weight, "weight")) || (rc = check_ptr(bias, "bias", false)) || (rc = check_ptr(out, "out")))
    return rc;
  if (!layer_ok(s, g, DCN_PHASE_LAYER_FORWARD)) {
    set_error("layer forward: shape / operand / flags not supported (dcn_path_name(s, DCN_PHASE_LAYER_FORWARD))");
    return DCN_ERR_UNSUPPORTED;
  }
  const size_t need = layer_workspace(g, s->operand, DCN_PHASE_LAYER_FORWARD);
  if (workspace_bytes < need) {
    set_error("layer forward workspace: have %zu bytes, need %zu", workspace_bytes, need);
    return DCN_ERR_WORKSPACE;
  }
  if ((rc = dcn_offset_conv_forward(s, x, offset_weight, offset_bias, offset, workspace, workspace_bytes, stream)))
    return rc;
  return umma_forward(g, s->operand, s->flags | DCN_FLAG_XT_STAGED, x, (const float*)offset, weight,
                      (const float*)bias, out, workspace, (cudaStream_t)stream);
}

// ---- chained inference: the producer's epilogue writes the consumer's staged input --------------------
// Geo of the producer with out_framed set from the consumer's staging layout, or an error.
static int chained_geo(const DcnShape* s, const DcnShape* consumer, Geo* g) {
  int rc = geo_or_error(s, g);
  if (rc) return rc;
  Geo gc;
  if ((rc = geo_or_error(consumer, &gc))) return rc;
  if (s->operand != DCN_OPERAND_FP32 || consumer->operand != DCN_OPERAND_FP32 || g->O > 256 ||
      gc.B != g->B || gc.C != g->O || gc.H != g->Ho || gc.W != g->Wo) {
    set_error("chained forward: consumer must take this layer's output (B=%d C=%d H=%d W=%d, fp32, O <= 256); got "
              "B=%d C=%d H=%d W=%d", g->B, g->O, g->Ho, g->Wo, gc.B, gc.C, gc.H, gc.W);
    return DCN_ERR_BAD_SHAPE;
  }
  if (!use_umma(s, *g, DCN_PHASE_FORWARD) || !use_umma(consumer, gc, DCN_PHASE_FORWARD)) {
    set_error("chained forward: both layers must run on the tensor path (dcn_path_name == \"umma\")");
    return DCN_ERR_UNSUPPORTED;
  }
  Tiling tc;
  if (!make_tiling(gc, &tc)) return DCN_ERR_UNSUPPORTED;
  g->out_framed = 1;
  g->out_G = gc.variant == DCN_VARIANT_TORCH ? tc.G : 0;
  g->out_Cs = gc.variant == DCN_VARIANT_TORCH ? tc.Cs : 0;
  return DCN_OK;
}

size_t dcn_staged_input_bytes(const DcnShape* s) {
  Geo g;
  if (geo_or_error(s, &g)) return 0;
  return align_up(sizeof(float) * (size_t)g.B * xt_image_stride(g), 1024);
}

int dcn_staged_input_clear(const DcnShape* s, void* workspace, void* stream) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  if ((rc = check_ptr(workspace, "workspace"))) return rc;
  DCN_CUDA_TRY(cudaMemsetAsync(workspace, 0, dcn_staged_input_bytes(s), (cudaStream_t)stream));
  return DCN_OK;
}

int dcn_layer_forward_chained(const DcnShape* s, const DcnShape* consumer, const void* x, const void* offset_weight,
                              const void* offset_bias, const void* weight, const void* bias, void* offset,
                              void* consumer_workspace, void* workspace, size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = chained_geo(s, consumer, &g);
  if (rc) return rc;
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(offset_weight, "offset_weight")) ||
      (rc = check_ptr(offset_bias, "offset_bias", false)) || (rc = check_ptr(weight, "weight")) ||
      (rc = check_ptr(bias, "bias", false)) || (rc = check_ptr(offset, "offset")) ||
      (rc = check_ptr(consumer_workspace, "consumer_workspace")) || (rc = check_ptr(workspace, "workspace")))
    return rc;
  if (!layer_ok(s, g, DCN_PHASE_LAYER_FORWARD)) {
    set_error("chained layer forward: shape / operand / flags not supported (dcn_path_name(s, DCN_PHASE_LAYER_FORWARD))");
    return DCN_ERR_UNSUPPORTED;
  }
  const size_t need = layer_workspace(g, s->operand, DCN_PHASE_LAYER_FORWARD);
  if (workspace_bytes < need) {
    set_error("chained layer forward workspace: have %zu bytes, need %zu", workspace_bytes, need);
    return DCN_ERR_WORKSPACE;
  }
  if ((rc = dcn_offset_conv_forward(s, x, offset_weight, offset_bias, offset, workspace, workspace_bytes, stream)))
    return rc;
  return umma_forward(g, s->operand, s->flags | DCN_FLAG_XT_STAGED, x, (const float*)offset, weight,
                      (const float*)bias, consumer_workspace, workspace, (cudaStream_t)stream);
}

static int staged_consumer(const DcnShape* consumer, Geo* g, Tiling* t) {
  int rc = geo_or_error(consumer, g);
  if (rc) return rc;
  if (consumer->operand != DCN_OPERAND_FP32 || !use_umma(consumer, *g, DCN_PHASE_FORWARD) || !make_tiling(*g, t)) {
    set_error("staged post-op: the consumer layer must run on the tensor path with fp32 operands");
    return DCN_ERR_UNSUPPORTED;
  }
  return DCN_OK;
}

int dcn_bn_relu_forward_staged(const DcnShape* consumer, int training, const void* x, const void* gamma, const void* beta,
                               void* running_mean, void* running_var, float momentum, float eps,
                               void* consumer_workspace, void* saved, void* workspace, size_t workspace_bytes,
                               void* stream) {
  Geo g;
  Tiling t;
  int rc = staged_consumer(consumer, &g, &t);
  if (rc) return rc;
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(consumer_workspace, "consumer_workspace")) ||
      (rc = check_ptr(saved, "saved")) || (rc = check_ptr(workspace, "workspace")))
    return rc;
  if (!training && (!running_mean || !running_var)) {
    set_error("staged post-op: eval mode needs running_mean / running_var");
    return DCN_ERR_NULL_POINTER;
  }
  if (workspace_bytes < bn_workspace_bytes(g.C)) {
    set_error("staged post-op workspace: have %zu bytes, need %zu", workspace_bytes, bn_workspace_bytes(g.C));
    return DCN_ERR_WORKSPACE;
  }
  return bn_relu_forward_staged(g, t, training, (const float*)x, (const float*)gamma, (const float*)beta,
                                (float*)running_mean, (float*)running_var, momentum, eps, (float*)consumer_workspace,
                                (float*)saved, workspace, (cudaStream_t)stream);
}

int dcn_bn_relu_backward_staged(const DcnShape* consumer, int training, const void* x, const void* grad_staged,
                                const void* saved, void* grad_x, void* grad_gamma, void* grad_beta, void* workspace,
                                size_t workspace_bytes, void* stream) {
  Geo g;
  Tiling t;
  int rc = staged_consumer(consumer, &g, &t);
  if (rc) return rc;
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(grad_staged, "grad_staged")) || (rc = check_ptr(saved, "saved")) ||
      (rc = check_ptr(grad_x, "grad_x", false)) || (rc = check_ptr(workspace, "workspace")))
    return rc;
  if (workspace_bytes < bn_workspace_bytes(g.C)) {
    set_error("staged post-op workspace: have %zu bytes, need %zu", workspace_bytes, bn_workspace_bytes(g.C));
    return DCN_ERR_WORKSPACE;
  }
  return bn_relu_backward_staged(g, t, training, (const float*)x, (const float*)grad_staged, (const float*)saved,
                                 (float*)grad_x, (float*)grad_gamma, (float*)grad_beta, workspace, (cudaStream_t)stream);
}

int dcn_layer_backward(const DcnShape* s, const void* x, const void* offset, const void* offset_weight,
                       const void* weight, const void* grad_out, void* grad_x, void* grad_offset_weight,
                       void* grad_offset_bias, void* grad_weight, void* grad_bias, void* workspace,
                       size_t workspace_bytes, void* stream) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  const bool want_gx = !(s->flags & DCN_FLAG_NO_GRAD_X);
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(offset, "offset")) ||
      (rc = check_ptr(offset_weight, "offset_weight")) || (rc = check_ptr(weight, "weight")) ||
      (rc = check_ptr(grad_out, "grad_out")) || (rc = check_ptr(grad_x, "grad_x", want_gx)) ||
      (rc = check_ptr(grad_offset_weight, "grad_offset_weight")) ||
      (rc = check_ptr(grad_offset_bias, "grad_offset_bias", false)) || (rc = check_ptr(grad_weight, "grad_weight")) ||
      (rc = check_ptr(grad_bias, "grad_bias", false)) || (rc = check_ptr(workspace, "workspace")))
    return rc;
  if (!layer_ok(s, g, DCN_PHASE_LAYER_BACKWARD)) {
    set_error("layer backward: shape / operand / flags not supported (dcn_path_name(s, DCN_PHASE_LAYER_BACKWARD))");
    return DCN_ERR_UNSUPPORTED;
  }
  const size_t need = layer_workspace(g, s->operand, DCN_PHASE_LAYER_BACKWARD);
  if (workspace_bytes < need) {
    set_error("layer backward workspace: have %zu bytes, need %zu", workspace_bytes, need);
    return DCN_ERR_WORKSPACE;
  }
  return umma_layer_backward(g, s->flags, x, (const float*)offset, (const float*)offset_weight, weight, grad_out,
                             (float*)grad_x, (float*)grad_offset_weight, (float*)grad_offset_bias,
                             (float*)grad_weight, (float*)grad_bias, workspace, need, (cudaStream_t)stream);
}

int dcn_debug_corners(const DcnShape* s, const void* offset, int32_t* y0, int32_t* x0, float* w4,
                      void* stream) {
  Geo g;
  int rc = geo_or_error(s, &g);
  if (rc) return rc;
  if ((rc = check_ptr(offset, "offset")) || (rc = check_ptr(y0, "y0")) ||
      (rc = check_ptr(x0, "x0")) || (rc = check_ptr(w4, "w4")))
    return rc;
  return launch_corners(g, (const float*)offset, y0, x0, w4, (cudaStream_t)stream);
}

size_t dcn_bn_workspace_bytes(int32_t C) { return C > 0 ? bn_workspace_bytes(C) : 0; }

int dcn_bn_relu_forward(int32_t B, int32_t C, int32_t HW, int32_t training, const void* x, const void* gamma,
                        const void* beta, void* running_mean, void* running_var, float momentum, float eps,
                        void* y, void* saved, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = bn_shape_ok(B, C, HW);
  if (rc) return rc;
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(y, "y")) || (rc = check_ptr(saved, "saved")) ||
      (rc = check_ptr(workspace, "workspace")))
    return rc;
  if (!training && (!running_mean || !running_var)) {
    set_error("eval-mode batch norm needs running_mean and running_var");
    return DCN_ERR_NULL_POINTER;
  }
  if (workspace_bytes < bn_workspace_bytes(C)) {
    set_error("batch-norm workspace: have %zu bytes, need %zu", workspace_bytes, bn_workspace_bytes(C));
    return DCN_ERR_WORKSPACE;
  }
  return bn_relu_forward(B, C, HW, training, (const float*)x, (const float*)gamma, (const float*)beta,
                         (float*)running_mean, (float*)running_var, momentum, eps, (float*)y, (float*)saved,
                         workspace, (cudaStream_t)stream);
}

int dcn_bn_relu_backward(int32_t B, int32_t C, int32_t HW, int32_t training, const void* x, const void* grad_y,
                         const void* saved, void* grad_x, void* grad_gamma, void* grad_beta, void* workspace,
                         size_t workspace_bytes, void* stream) {
  int rc = bn_shape_ok(B, C, HW);
  if (rc) return rc;
  if ((rc = check_ptr(x, "x")) || (rc = check_ptr(grad_y, "grad_y")) || (rc = check_ptr(saved, "saved")) ||
      (rc = check_ptr(workspace, "workspace")))
    return rc;
  if (workspace_bytes < bn_workspace_bytes(C)) {
    set_error("batch-norm workspace: have %zu bytes, need %zu", workspace_bytes, bn_workspace_bytes(C));
    return DCN_ERR_WORKSPACE;
  }
  return bn_relu_backward(B, C, HW, training, (const float*)x, (const float*)grad_y, (const float*)saved,
                          (float*)grad_x, (float*)grad_gamma, (float*)grad_beta, workspace, (cudaStream_t)stream);
}

}  // extern "C"
