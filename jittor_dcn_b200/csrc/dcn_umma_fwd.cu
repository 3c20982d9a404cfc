// dcn_umma_fwd.cu — forward pass as ONE warp-specialised implicit GEMM on the 5th-gen tensor
// cores (tcgen05.mma, accumulators in TMEM).
//
// Replaces, per 128-row tile and without ever writing columns to HBM:
//   coordinates + bilinear sampling   deform_conv.py:62-68,30-54 / train.py:102-127
//   column layout                     deform_conv.py:72-73       / train.py:129-131
//   GEMM + bias + NCHW store          deform_conv.py:74-81       / train.py:133-140
//
// Warp roles (832 threads, 1 CTA per SM, persistent over tiles; the gather is issue/latency
// bound, so it gets most of the warps):
//   warps 0-3   epilogue : tcgen05.ld accumulator -> + bias -> out[B,O,Ho,Wo]
//   warp  4     MMA      : one lane issues tcgen05.mma (bf16 hi/lo split, 3 MMAs per K step)
//   warp  5     B loader : one lane streams the pre-tiled weight images with cp.async.bulk
//   warps 6-9   plan     : offsets -> bit-exact coordinate chain -> branch-free gather entries
//                          (corner offsets + masked weights) in a 2-deep smem ring, one K block
//                          ahead of the gather warps
//   warps 10-25 gather   : 4x LDG.128 per entry from the channels-last input -> blend ->
//                          bf16 hi/lo -> swizzled smem A images
// Pipelines (all mbarrier based): plan ring (plan -> gather -> back), smem stages
// (gather + loader -> MMA -> back), TMEM accumulators (MMA -> epilogue -> back; 2 buffers).
//
// fp32 parity: every fp32 operand v is split as hi = bf16(v), lo = bf16(v - hi) and the product
// is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation (relative error ~5e-6, measured).
#include <cstdlib>

#include "dcn_umma.h"
#include "dcn_umma_common.cuh"

namespace dcn {

using namespace ptx;

constexpr int kEpiWarps = 4, kPlanWarps = 4, kProdWarps = 16;
constexpr int kPlanThreads = kPlanWarps * 32;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kFirstPlanWarp = kEpiWarps + 2, kFirstProdWarp = kFirstPlanWarp + kPlanWarps;
constexpr int kFwdThreads = (kFirstProdWarp + kProdWarps) * 32;  // 832
constexpr int kPlanPerThread = 4;                                // kPlanMax / kPlanThreads
constexpr int kItems = 128 * 16 / kProdThreads;                  // float4 items per thread and K block (4)
constexpr int kPlanMax = 512;                                    // plan entries per K block (32 B each)
constexpr uint32_t kATile = 128 * 64 * 2;                        // one bf16 A image (16 KB)
constexpr uint32_t kAMnLbo = 1024, kAMnSbo = 2048;               // MN-major A: atom strides
constexpr int kMaxStages = 4;

struct FwdParams {
  Geo g;
  Tiling t;
  const float* xt;
  const float* off;
  const uint8_t* wtiles;
  const float* bias;
  float* out;
  int stages;
  int dbg;             // DCN_FWD_DBG timing experiments (results invalid when non-zero)
  int plan_cap;        // plan entries per buffer (n_ent rounded up to 256)
  uint32_t b_tile;     // bytes of one bf16 B image = O*128
  uint32_t stage_bytes;
  uint32_t tmem_cols;  // 2*O rounded to a power of two >= 32
};

struct TileRowInfo {  // what a Torch-layout tile row needs
  int b, r0, chunk, valid;
};

__device__ __forceinline__ TileRowInfo decode_inst(const Tiling& t, int inst) {
  TileRowInfo ri;
  ri.valid = inst < t.num_inst;
  uint32_t bc, r0, b, ch;
  t.divR.divmod((uint32_t)(ri.valid ? inst : 0), bc, r0);
  t.divChunks.divmod(bc, b, ch);
  ri.b = (int)b;
  ri.r0 = (int)r0;
  ri.chunk = (int)ch;
  return ri;
}

// One plan entry in the making: the index math is done and the two offset loads are in
// flight (issued one K block ahead so that their latency hides behind the gather phase).
struct PlanWork {
  float ox, oy;
  int h, w, chan_base, valid;
};

template <int VARIANT>
__device__ __forceinline__ void plan_prepare(const Geo& g, const Tiling& t, const float* __restrict__ off,
                                             int tile, int kb, int e, PlanWork& pw) {
  pw.valid = 0;
  pw.ox = pw.oy = 0.f;
  pw.h = pw.w = pw.chan_base = 0;
  int b, p, n;
  if (VARIANT == DCN_VARIANT_TORCH) {
    const int il = e >> 6, kk = e & 63, j = kb * 64 + kk;
    const TileRowInfo ri = decode_inst(t, tile * t.Rt + il);
    if (!ri.valid || j >= g.K) return;
    uint32_t cb, q, pp, nn;
    t.divP.divmod((uint32_t)(ri.r0 * g.K + j), cb, q);
    t.divN.divmod(q, pp, nn);
    b = ri.b;
    p = (int)pp;
    n = (int)nn;
    pw.chan_base = (int)cb * t.G + ri.chunk * t.Gt;
  } else {
    const int tl = e >> 7, m = e & 127;
    b = tile / t.pix_blocks;
    p = (tile - b * t.pix_blocks) * 128 + m;
    n = (kb * 64) / g.C + tl;
    if (p >= g.HW || n >= g.N) return;
  }
  uint32_t h, w;
  t.divWo.divmod((uint32_t)p, h, w);
  pw.h = (int)h;
  pw.w = (int)w;
  const float* ob = off + (size_t)b * 2 * g.N * g.HW;
  pw.ox = __ldg(ob + (size_t)n * g.HW + p);
  pw.oy = __ldg(ob + (size_t)(g.N + n) * g.HW + p);
  pw.valid = 1;
}

// coordinate chain (bit-exact, dcn_common.cuh:tap_of) -> branch-free gather entry
__device__ __forceinline__ PlanEntry plan_finish(const Geo& g, const PlanWork& pw) {
  PlanEntry e;
  const int pad = g.H * g.W * g.C + pw.chan_base;  // the zero pixel
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    e.off[k] = pad;
    e.w[k] = 0.f;
  }
  if (pw.valid) {
    const Tap tp = tap_of(g, pw.h, pw.w, pw.ox, pw.oy);
    const unsigned m = corner_mask(tp, g.H, g.W);
    if (m) {
      float cw[4];
      corner_weights(tp, cw);
      const int base = (tp.y0 * g.W + tp.x0) * g.C + pw.chan_base;  // only used for valid corners
      if (m & 1u) { e.off[0] = base;                     e.w[0] = cw[0]; }
      if (m & 2u) { e.off[1] = base + g.C;               e.w[1] = cw[1]; }
      if (m & 4u) { e.off[2] = base + g.W * g.C;         e.w[2] = cw[2]; }
      if (m & 8u) { e.off[3] = base + g.W * g.C + g.C;   e.w[3] = cw[3]; }
    }
  }
  return e;
}

__device__ __forceinline__ float4 blend4(const float4 v[4], const PlanEntry& e) {
  float4 r;
  r.x = fmaf(v[3].x, e.w[3], fmaf(v[2].x, e.w[2], fmaf(v[1].x, e.w[1], v[0].x * e.w[0])));
  r.y = fmaf(v[3].y, e.w[3], fmaf(v[2].y, e.w[2], fmaf(v[1].y, e.w[1], v[0].y * e.w[0])));
  r.z = fmaf(v[3].z, e.w[3], fmaf(v[2].z, e.w[2], fmaf(v[1].z, e.w[1], v[0].z * e.w[0])));
  r.w = fmaf(v[3].w, e.w[3], fmaf(v[2].w, e.w[2], fmaf(v[1].w, e.w[1], v[0].w * e.w[0])));
  return r;
}

// float4 -> 4 bf16 hi + 4 bf16 lo, each packed in 8 bytes
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  split_pair(v.x, v.y, hi.x, lo.x);
  split_pair(v.z, v.w, hi.y, lo.y);
}

template <int VARIANT>
__global__ void __launch_bounds__(kFwdThreads, 1) umma_fwd_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic (a uintptr_t round trip would lose the shared
  // address space and turn every LDS/STS below into a slow generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const Geo& g = P.g;
  const Tiling& t = P.t;
  // carve-up: [stages x (A_hi | A_lo | B_hi | B_lo)] [plan x2] [barriers]
  uint8_t* stage_base = smem;
  PlanEntry* plan = reinterpret_cast<PlanEntry*>(smem + (size_t)P.stages * P.stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(plan + 2 * P.plan_cap);
  uint64_t* full = bars;                   // [stages]
  uint64_t* empty = bars + kMaxStages;     // [stages]
  uint64_t* tfull = bars + 2 * kMaxStages; // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint64_t* pfull = tempty + 2;            // [2] plan ring
  uint64_t* pempty = pfull + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int O = g.O;

  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], kProdWarps + 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps);
      mbar_init(&pfull[a], kPlanWarps);
      mbar_init(&pempty[a], kProdWarps);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ================================================================ epilogue
    uint32_t acc_phase = 0;
    int acc = 0;
    for (int tile = blockIdx.x; tile < t.num_tiles; tile += gridDim.x) {
      const int m = warp * 32 + lane;  // TMEM lane == tile row
      size_t out_off = 0;
      bool valid;
      if (VARIANT == DCN_VARIANT_TORCH) {
        const int quad = m / (4 * t.Rt), il = (m >> 2) % t.Rt, chq = m & 3;
        const TileRowInfo ri = decode_inst(t, tile * t.Rt + il);
        valid = ri.valid;
        const int i = ri.chunk * t.Gt + 4 * quad + chq;
        out_off = (size_t)ri.b * O * g.HW + (size_t)(ri.r0 + i * t.R);
      } else {
        const int b = tile / t.pix_blocks, p = (tile - b * t.pix_blocks) * 128 + m;
        valid = p < g.HW;
        out_off = (size_t)b * O * g.HW + p;
      }
      mbar_wait_relaxed(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * O);
      for (int c0 = 0; c0 < O; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (valid && !(P.dbg & 16)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float bv = P.bias ? __ldg(P.bias + c0 + i) : 0.f;
            P.out[out_off + (size_t)(c0 + i) * g.HW] = v[i] + bv;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp == kEpiWarps) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, O, VARIANT == DCN_VARIANT_TORCH, false);
      int s = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < t.num_tiles; tile += gridDim.x) {
        mbar_wait_relaxed(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * O);
        for (int kb = 0; kb < t.KB; ++kb) {
          mbar_wait_relaxed(&full[s], phase, 64);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(stage_base + (size_t)s * P.stage_bytes);
          const uint32_t a_lo = a_hi + kATile;
          const uint32_t b_hi = a_hi + 2 * kATile;
          const uint32_t b_lo = b_hi + P.b_tile;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            if (P.dbg & 8) break;
            uint64_t dah, dal;
            if (VARIANT == DCN_VARIANT_TORCH) {
              dah = make_sdesc_sw128(a_hi + k4 * 2 * kAMnSbo, kAMnLbo, kAMnSbo);
              dal = make_sdesc_sw128(a_lo + k4 * 2 * kAMnSbo, kAMnLbo, kAMnSbo);
            } else {
              dah = make_sdesc_sw128(a_hi + k4 * 32, 16, 1024);
              dal = make_sdesc_sw128(a_lo + k4 * 32, 16, 1024);
            }
            const uint64_t dbh = make_sdesc_sw128(b_hi + k4 * 32, 16, 1024);
            const uint64_t dbl = make_sdesc_sw128(b_lo + k4 * 32, 16, 1024);
            umma_bf16(d_tmem, dah, dbh, idesc, (kb | k4) ? 1u : 0u);
            umma_bf16(d_tmem, dah, dbl, idesc, 1u);
            umma_bf16(d_tmem, dal, dbh, idesc, 1u);
          }
          umma_commit(&empty[s]);  // stage reusable once these MMAs have read it
          if (kb == t.KB - 1) umma_commit(&tfull[acc]);
          if (++s == P.stages) {
            s = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ================================================================ weight (B) loader
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < t.num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < t.KB; ++kb) {
          mbar_wait_relaxed(&empty[s], phase ^ 1);
          uint8_t* dst = stage_base + (size_t)s * P.stage_bytes + 2 * kATile;
          mbar_arrive_expect_tx(&full[s], 2 * P.b_tile);
          bulk_g2s(dst, P.wtiles + (size_t)kb * 2 * P.b_tile, 2 * P.b_tile, &full[s]);
          if (++s == P.stages) {
            s = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp < kFirstProdWarp) {
    // ================================================================ plan warps
    const int pt = tid - kFirstPlanWarp * 32;  // 0..63
    const int n_ent = VARIANT == DCN_VARIANT_TORCH ? t.Rt * 64 : 128 * t.taps_per_kb;
    int pbuf = 0;
    uint32_t pphase = 0;
    PlanWork pw[kPlanPerThread];
    if ((int)blockIdx.x < t.num_tiles) {
#pragma unroll
      for (int u = 0; u < kPlanPerThread; ++u)
        if (pt + u * kPlanThreads < n_ent)
          plan_prepare<VARIANT>(g, t, P.off, blockIdx.x, 0, pt + u * kPlanThreads, pw[u]);
    }
    for (int tile = blockIdx.x; tile < t.num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < t.KB; ++kb) {
        PlanEntry* pl = plan + pbuf * P.plan_cap;
        mbar_wait_relaxed(&pempty[pbuf], pphase ^ 1, 64);
        // finish this K block's entries (their offset loads were issued one block ago) ...
#pragma unroll
        for (int u = 0; u < kPlanPerThread; ++u)
          if (pt + u * kPlanThreads < n_ent) {
            if (P.dbg & 2) {
              PlanEntry pe;
              pe.off[0] = pe.off[1] = pe.off[2] = pe.off[3] = (pt & 63) * g.C;
              pe.w[0] = pe.w[1] = pe.w[2] = pe.w[3] = 0.25f;
              pl[pt + u * kPlanThreads] = pe;
            } else {
              pl[pt + u * kPlanThreads] = plan_finish(g, pw[u]);
            }
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pfull[pbuf]);
        // ... and start the next block's
        int ntile = tile, nkb = kb + 1;
        if (nkb == t.KB) {
          nkb = 0;
          ntile = tile + gridDim.x;
        }
        if (ntile < t.num_tiles && !(P.dbg & 2)) {
#pragma unroll
          for (int u = 0; u < kPlanPerThread; ++u)
            if (pt + u * kPlanThreads < n_ent)
              plan_prepare<VARIANT>(g, t, P.off, ntile, nkb, pt + u * kPlanThreads, pw[u]);
        }
        pbuf ^= 1;
        if (pbuf == 0) pphase ^= 1;
      }
    }
  } else {
    // ================================================================ gather warps
    // kItems (4) float4 items per thread and K block, software pipelined one item deep: the 4
    // LDG.128 of item i+1 (possibly the first item of the NEXT K block) are issued before item
    // i is blended, so every warp always has gathers in flight.
    const int pt = tid - kFirstProdWarp * 32;  // 0..511
    const size_t img_stride = xt_image_stride(g);
    // `quads` lanes share one sampling point; a "pair" is one (class instance, column) of the
    // Torch tile or one row of the Jittor tile
    const int quads = VARIANT == DCN_VARIANT_TORCH ? (t.Gt >> 2) : 16;
    const int quad = pt % quads, slot = pt / quads;
    const int pairs_per_pass = kProdThreads / quads;
    int ent_idx[kItems], item_il[kItems];
    uint32_t st_off[kItems];
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int pair = slot + it * pairs_per_pass;
      if (VARIANT == DCN_VARIANT_TORCH) {
        const int il = pair >> 6, kk = pair & 63;
        item_il[it] = il;
        ent_idx[it] = pair;  // = il * 64 + kk
        st_off[it] = mnmajor_sw128_off(quad * (4 * t.Rt) + il * 4, kk, kAMnLbo, kAMnSbo);
      } else {
        item_il[it] = 0;
        ent_idx[it] = pair;  // row m
        st_off[it] = kmajor_sw128_off(pair, quad * 4);
      }
    }
    // position of one K block in the three rings it touches
    struct Pos {
      int tile, kb, s, pbuf;
      uint32_t phase, pphase;
    };
    auto advance = [&](Pos& q) {
      if (++q.kb == t.KB) {
        q.kb = 0;
        q.tile += gridDim.x;
      }
      if (++q.s == P.stages) {
        q.s = 0;
        q.phase ^= 1;
      }
      q.pbuf ^= 1;
      if (q.pbuf == 0) q.pphase ^= 1;
    };
    const float* img[kItems];  // load side: image base (+ channel quad) of each item's rows
    auto set_images = [&](int tile) {
#pragma unroll
      for (int it = 0; it < kItems; ++it) {
        if (VARIANT == DCN_VARIANT_TORCH)
          img[it] = P.xt + (size_t)decode_inst(t, tile * t.Rt + item_il[it]).b * img_stride + quad * 4;
        else
          img[it] = P.xt + (size_t)(tile / t.pix_blocks) * img_stride;
      }
    };
    // Jittor layout: channel offset / tap slot of this thread's 4 columns inside K block kb
    auto jit_cols = [&](int kb, int& c, int& tl, bool& ok) {
      const int j = kb * 64 + quad * 4;
      const int n = j / g.C;
      c = j - n * g.C;
      tl = n - (kb * 64) / g.C;
      ok = j < g.K;
    };
    float4 v[2][4];  // [buffer][corner]
    auto issue = [&](const Pos& q, int it, int buf) {
      const PlanEntry* pl = plan + q.pbuf * P.plan_cap;
      int jc = 0, tl = 0;
      bool ok = true;
      if (VARIANT != DCN_VARIANT_TORCH) {
        jit_cols(q.kb, jc, tl, ok);
        pl += tl * 128;
      }
      const int4 off = *reinterpret_cast<const int4*>(pl[ent_idx[it]].off);
      const float* base = img[it] + jc;
      if (P.dbg & 1) {
        v[buf][0] = v[buf][1] = v[buf][2] = v[buf][3] =
            make_float4((float)off.x, (float)off.y, (float)off.z, (float)off.w);
      } else {
        v[buf][0] = __ldg(reinterpret_cast<const float4*>(base + off.x));
        v[buf][1] = __ldg(reinterpret_cast<const float4*>(base + off.y));
        v[buf][2] = __ldg(reinterpret_cast<const float4*>(base + off.z));
        v[buf][3] = __ldg(reinterpret_cast<const float4*>(base + off.w));
      }
    };
    Pos cur{(int)blockIdx.x, 0, 0, 0, 0u, 0u};
    if (cur.tile < t.num_tiles) {
      set_images(cur.tile);
      mbar_wait(&pfull[cur.pbuf], cur.pphase);
      mbar_wait(&empty[cur.s], cur.phase ^ 1);
      issue(cur, 0, 0);
    }
    while (cur.tile < t.num_tiles) {
      Pos nxt = cur;
      advance(nxt);
      const PlanEntry* pl = plan + cur.pbuf * P.plan_cap;
      bool col_ok = true;
      if (VARIANT != DCN_VARIANT_TORCH) {
        int jc, tl;
        jit_cols(cur.kb, jc, tl, col_ok);
        pl += tl * 128;
      }
      uint8_t* a_hi = stage_base + (size_t)cur.s * P.stage_bytes;
      uint8_t* a_lo = a_hi + kATile;
#pragma unroll
      for (int it = 0; it < kItems; ++it) {
        // keep the L1 fed: next item first
        if (it < kItems - 1) {
          issue(cur, it + 1, (it + 1) & 1);
        } else if (nxt.tile < t.num_tiles) {
          if (nxt.tile != cur.tile) set_images(nxt.tile);
          mbar_wait(&pfull[nxt.pbuf], nxt.pphase);
          mbar_wait(&empty[nxt.s], nxt.phase ^ 1);
          issue(nxt, 0, 0);
        }
        float4 w = *reinterpret_cast<const float4*>(pl[ent_idx[it]].w);
        if (VARIANT != DCN_VARIANT_TORCH && !col_ok) w = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* c4 = v[it & 1];
        float4 r;
        r.x = fmaf(c4[3].x, w.w, fmaf(c4[2].x, w.z, fmaf(c4[1].x, w.y, c4[0].x * w.x)));
        r.y = fmaf(c4[3].y, w.w, fmaf(c4[2].y, w.z, fmaf(c4[1].y, w.y, c4[0].y * w.x)));
        r.z = fmaf(c4[3].z, w.w, fmaf(c4[2].z, w.z, fmaf(c4[1].z, w.y, c4[0].z * w.x)));
        r.w = fmaf(c4[3].w, w.w, fmaf(c4[2].w, w.z, fmaf(c4[1].w, w.y, c4[0].w * w.x)));
        uint2 hi, lo;
        split4(r, hi, lo);
        if (!(P.dbg & 4) || hi.x == 0x12345678u) {
          *reinterpret_cast<uint2*>(a_hi + st_off[it]) = hi;
          *reinterpret_cast<uint2*>(a_lo + st_off[it]) = lo;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full[cur.s]);
        mbar_arrive(&pempty[cur.pbuf]);
      }
      cur = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(P.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------- host side
static uint32_t pow2_cols(int cols) {
  uint32_t c = 32;
  while ((int)c < cols) c <<= 1;
  return c;
}

bool umma_fwd_supported(const Geo& g, int operand) {
  if (operand != DCN_OPERAND_FP32) return false;
  if (g.O % 16 || g.O < 16 || g.O > 256) return false;
  Tiling t;
  if (!make_tiling(g, &t)) return false;
  if (g.variant == DCN_VARIANT_TORCH && t.Rt * 64 > kPlanMax) return false;
  static_assert(kPlanMax == kPlanPerThread * kPlanThreads, "plan slots");
  if (g.variant == DCN_VARIANT_JITTOR && 128 * t.taps_per_kb > kPlanMax) return false;
  return true;
}

size_t umma_fwd_workspace(const Geo& g) {
  Tiling t;
  make_tiling(g, &t);
  const size_t xt = align_up(sizeof(float) * (size_t)g.B * xt_image_stride(g), 1024);
  const size_t wt = align_up((size_t)t.KB * 2 * g.O * 128, 1024);
  return xt + wt;
}

int umma_forward_fp32(const Geo& g, const float* x, const float* off, const float* wt, const float* bias,
                      float* out, void* workspace, cudaStream_t st) {
  FwdParams P;
  P.g = g;
  if (!make_tiling(g, &P.t)) {
    set_error("umma forward: shape not tileable");
    return DCN_ERR_UNSUPPORTED;
  }
  float* xt = (float*)workspace;
  uint8_t* wtiles = (uint8_t*)workspace + align_up(sizeof(float) * (size_t)g.B * xt_image_stride(g), 1024);
  int rc;
  if ((rc = launch_nchw_to_nhwc(g, P.t, x, xt, st))) return rc;
  if ((rc = launch_weight_tiles_fwd(g, P.t, wt, wtiles, st))) return rc;
  P.xt = xt;
  P.off = off;
  P.wtiles = wtiles;
  P.bias = bias;
  P.out = out;
  P.b_tile = (uint32_t)g.O * 128;
  P.stage_bytes = 2 * kATile + 2 * P.b_tile;
  const int n_ent = g.variant == DCN_VARIANT_TORCH ? P.t.Rt * 64 : 128 * P.t.taps_per_kb;
  P.plan_cap = (n_ent + 255) / 256 * 256;
  const size_t fixed = 2 * (size_t)P.plan_cap * sizeof(PlanEntry) + 256 + 1024;  // plan ring, barriers, align slack
  int stages = (int)((227 * 1024 - fixed) / P.stage_bytes);
  // Two stages are enough to overlap the (fast) MMAs with the (slow) gather, and every KB of
  // shared memory not taken is L1 for the gather's footprint (L1 + smem share 256 KB).
  int want = 2;
  P.dbg = 0;
  if (const char* e = getenv("DCN_FWD_DBG")) P.dbg = atoi(e);
  if (const char* e = getenv("DCN_FWD_STAGES")) want = atoi(e);
  if (want < 2) want = 2;
  if (want > kMaxStages) want = kMaxStages;
  P.stages = stages > want ? want : stages;
  if (P.stages < 2) {
    set_error("umma forward: not enough shared memory for 2 stages");
    return DCN_ERR_UNSUPPORTED;
  }
  P.tmem_cols = pow2_cols(2 * g.O);
  const size_t smem = (size_t)P.stages * P.stage_bytes + fixed;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = P.t.num_tiles < sms ? P.t.num_tiles : sms;
  KernelScope scope("umma_fwd_kernel", st);
  if (g.variant == DCN_VARIANT_TORCH) {
    DCN_CUDA_TRY(cudaFuncSetAttribute(umma_fwd_kernel<DCN_VARIANT_TORCH>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_fwd_kernel<DCN_VARIANT_TORCH><<<grid, kFwdThreads, smem, st>>>(P);
  } else {
    DCN_CUDA_TRY(cudaFuncSetAttribute(umma_fwd_kernel<DCN_VARIANT_JITTOR>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_fwd_kernel<DCN_VARIANT_JITTOR><<<grid, kFwdThreads, smem, st>>>(P);
  }
  DCN_KERNEL_CHECK("umma_fwd_kernel");
  return DCN_OK;
}

}  // namespace dcn
