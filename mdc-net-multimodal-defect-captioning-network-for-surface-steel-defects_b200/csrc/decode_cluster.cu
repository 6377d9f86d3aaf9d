// decode_cluster.cu -- the whole autoregressive decode loop as ONE persistent, cluster-cooperative kernel
// (north_star (c)+(d): paged bf16 self-KV, HBM/L2-resident cross-K/V, LayerNorm + residual + token select fused;
//  "no collective inside the decode loop" -- here not even a kernel boundary).
//
// Decomposition (model width 256, 8 heads x 32, FFN 2048 -- the configuration inference_p.py:126-129 builds):
//   * a thread-block CLUSTER of 8 CTAs owns G images (G = ceil(B / #clusters) <= 16) for the entire loop;
//     clusters never talk to each other, so there is no grid-wide synchronisation at all.
//   * inside a cluster CTA r owns attention head r and 1/8 of every projection:
//       in-proj rows of head r (q,k,v), out-proj / cross-q / cross-out rows [32r,32r+32), FFN1 hidden units
//       [256r,256r+256), FFN2 as a K-split over the same hidden units, vocab rows [40r,40r+40).
//     Activations (a few KB) are exchanged through DISTRIBUTED SHARED MEMORY with cluster barriers; weights, the
//     paged self-KV cache and the resident cross-K/V STREAM through a 4-stage cp.async ring (one 33 KB stage =
//     64 weight rows x 256 k, or one K / V panel), so HBM/L2 traffic stays in flight across phase boundaries.
//   * projections run on the tensor cores: mma.sync m16n8k16 bf16 with the G <= 16 images as the M dimension.
//     Activations are split hi+lo into two bf16 operands (x = hi + lo, two MMAs), so the bf16 WEIGHTS are the only
//     rounding on the fast path -- the same numerics as the unfused fp32-activation kernels in decode.cu.
//   * LayerNorm+residual are fused into the DSMEM gathers; the head's logits go to a small global buffer, CTA r
//     runs the greedy / top-k / top-p select for images r, r+8 and publishes the token; one more cluster barrier
//     and the next step's embedding gather starts.  Weight prefetch runs across steps.
// The generic kernels in decode.cu remain the path for the fp32 token-exact mode, teacher forcing and other
// geometries; mdc_decode_steps picks this kernel when the geometry matches.
#include "common.cuh"
#include "select.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

namespace cg = cooperative_groups;
using namespace mdcsel;

namespace {

constexpr int CS = 8;              // cluster size == heads
constexpr int DM = 256;            // model width
constexpr int HD = 32;             // head width
constexpr int FFN = 2048;
constexpr int FS = FFN / CS;       // hidden units per CTA (256)
constexpr int NT = 256;            // threads per CTA
constexpr int GM = 16;             // max images per cluster (the MMA M dimension)
constexpr int PITCH = DM + 8;      // bf16 elements per padded smem row (528 B: conflict-free fragment loads)
constexpr int STAGE_ROWS = 64;
constexpr int STAGE_BYTES = STAGE_ROWS * PITCH * 2;   // 33792
constexpr int NS = 4;              // ring stages
constexpr int KVP = 40;            // bf16 elements per padded K/V row in a stage (80 B)
constexpr int VSL = 40;            // vocab rows per CTA (5 n-tiles); 8*40 = 320 >= V
constexpr int SCR = 264;           // floats per warp score row: <= 256 keys + own key + normaliser

struct ClusterParams {
  // weights (bf16 [N,K] row-major unless noted; f32 for biases / norms / tables)
  const bf16* w_in[8]; const float* b_in[8]; const bf16* w_so[8]; const float* b_so[8];
  const float* ln1w[8]; const float* ln1b[8];
  const bf16* w_ca[8]; const float* b_ca[8]; const bf16* w_co[8]; const float* b_co[8];
  const float* ln2w[8]; const float* ln2b[8];
  const bf16* w_f1[8]; const float* b_f1[8]; const bf16* w_f2[8]; const float* b_f2[8];
  const float* ln3w[8]; const float* ln3b[8];
  const float* emb; const float* pos; const bf16* w_out; const float* b_out;
  int layers, vocab, S, pad_idx;
  // batch state
  int B, G;
  int32_t* tokens; int tokens_ld;
  bf16* kv_pool; const int32_t* page_table; int pages_per_seq, PT;
  const bf16* cross_kv;                    // [layer][B*S][2*DM]
  float* step_logits;                      // [B][vocab] scratch (L2)
  float* logits_out; int64_t logits_img_stride; int logits_row_offset;
  float* confs; int confs_ld;
  const float* uniforms; int uniforms_ld; int top_k; float top_p; int forced;
  int t_begin, t_end, maxT;                // maxT: padded key capacity per image in a self-KV stage
  int ips;                                 // images per self-KV stage
};

// ---- small helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ void cp16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_store(bf16* hi, bf16* lo, int idx, float x) {
  bf16 h = __float2bfloat16_rn(x);
  hi[idx] = h; lo[idx] = __float2bfloat16_rn(x - __bfloat162float(h));
}

struct Smem {
  uint8_t* ring;        // NS * STAGE_BYTES
  bf16 *a_hi, *a_lo;    // [GM][PITCH]  A operand of the width-256 projections
  bf16 *f_hi, *f_lo;    // [GM][PITCH]  A operand of FFN2 (own hidden slice)
  float* xres;          // [GM][DM]     residual stream
  float* qkv;           // [GM][96]     own head's q|k|v
  float* qc;            // [GM][32]
  float* oslice;        // [GM][32]     DSMEM-exposed: own head's attention output
  float* yslice;        // [GM][32]     DSMEM-exposed: own 32-column slice of a projection
  float* ypart;         // [GM][DM]     DSMEM-exposed: FFN2 partial sums
  float* scores;        // [8 warps][SCR]   (also the select scratch)
  int* pages;           // [GM][32]
  uint8_t* padflag;     // [GM][256]
};

__device__ __forceinline__ Smem carve(uint8_t* base) {
  Smem s; uint8_t* p = base;
  s.ring = p; p += NS * STAGE_BYTES;
  s.a_hi = (bf16*)p; p += GM * PITCH * 2; s.a_lo = (bf16*)p; p += GM * PITCH * 2;
  s.f_hi = (bf16*)p; p += GM * PITCH * 2; s.f_lo = (bf16*)p; p += GM * PITCH * 2;
  s.xres = (float*)p; p += GM * DM * 4;
  s.qkv = (float*)p; p += GM * 96 * 4;
  s.qc = (float*)p; p += GM * 32 * 4;
  s.oslice = (float*)p; p += GM * 32 * 4;
  s.yslice = (float*)p; p += GM * 32 * 4;
  s.ypart = (float*)p; p += GM * DM * 4;
  s.scores = (float*)p; p += 8 * SCR * 4;
  s.pages = (int*)p; p += GM * 32 * 4;
  s.padflag = p; p += GM * 256;
  return s;
}
constexpr int SMEM_BYTES = NS * STAGE_BYTES + 4 * GM * PITCH * 2 + GM * DM * 4 + GM * 96 * 4 + 3 * GM * 32 * 4 + GM * DM * 4 +
                           8 * SCR * 4 + GM * 32 * 4 + GM * 256;

// ---- the stage schedule -------------------------------------------------------------------------------
// per layer: 0,1 in-proj (q+k rows | v rows); (selfK, selfV) x nS; 1 self-out; 1 cross-q; (crossK, crossV) x nC;
//            1 cross-out; 4 FFN1; 4 FFN2.   After the last layer: 1 head stage.
struct Sched {
  int nS, nC, spl, sps;   // self stages, cross stages, stages per layer, stages per step
  __device__ Sched(int G, int ips) {
    nS = (G + ips - 1) / ips; nC = (G + 1) / 2;
    spl = 2 + 2 * nS + 2 + 2 * nC + 1 + 8;
    sps = 0;
  }
};

enum { K_INA = 0, K_INB, K_SELFK, K_SELFV, K_SOUT, K_CQ, K_CROSSK, K_CROSSV, K_COUT, K_F1, K_F2, K_HEAD };

struct StageId { int kind, sub, layer, t; };

__device__ __forceinline__ StageId decode_stage(const ClusterParams& P, const Sched& sc, int64_t idx) {
  const int sps = P.layers * sc.spl + 1;
  StageId id; id.t = P.t_begin + (int)(idx / sps);
  int j = (int)(idx % sps);
  if (j == sps - 1) { id.kind = K_HEAD; id.sub = 0; id.layer = 0; return id; }
  id.layer = j / sc.spl; j %= sc.spl;
  if (j < 2) { id.kind = j == 0 ? K_INA : K_INB; id.sub = 0; return id; } j -= 2;
  if (j < 2 * sc.nS) { id.kind = (j & 1) ? K_SELFV : K_SELFK; id.sub = j >> 1; return id; } j -= 2 * sc.nS;
  if (j == 0) { id.kind = K_SOUT; id.sub = 0; return id; } j -= 1;
  if (j == 0) { id.kind = K_CQ; id.sub = 0; return id; } j -= 1;
  if (j < 2 * sc.nC) { id.kind = (j & 1) ? K_CROSSV : K_CROSSK; id.sub = j >> 1; return id; } j -= 2 * sc.nC;
  if (j == 0) { id.kind = K_COUT; id.sub = 0; return id; } j -= 1;
  if (j < 4) { id.kind = K_F1; id.sub = j; return id; } j -= 4;
  id.kind = K_F2; id.sub = j; return id;
}

// copy `nrows` weight rows (256 bf16 each, source pitch `spitch` elements, first row `row0`, k offset `koff`) to stage rows [dst0, ...)
__device__ __forceinline__ void issue_rows(uint8_t* stage, int dst0, const bf16* W, int row0, int nrows, int spitch, int koff) {
  for (int c = threadIdx.x; c < nrows * 32; c += NT) {
    const int r = c >> 5, ch = c & 31;
    cp16(stage + (size_t)(dst0 + r) * PITCH * 2 + ch * 16, W + (int64_t)(row0 + r) * spitch + koff + ch * 8);
  }
}

__device__ void issue_stage(const ClusterParams& P, const Sched& sc, const Smem& sm, int64_t idx, int64_t total, int rank, int img0, int G) {
  if (idx < total) {
    uint8_t* stage = sm.ring + (size_t)(idx % NS) * STAGE_BYTES;
    const StageId id = decode_stage(P, sc, idx);
    const int l = id.layer;
    switch (id.kind) {
      case K_INA:
        issue_rows(stage, 0, P.w_in[l], rank * HD, 32, DM, 0);                 // q rows of head `rank`
        issue_rows(stage, 32, P.w_in[l], DM + rank * HD, 32, DM, 0);           // k rows
        break;
      case K_INB: issue_rows(stage, 0, P.w_in[l], 2 * DM + rank * HD, 32, DM, 0); break;   // v rows
      case K_SOUT: issue_rows(stage, 0, P.w_so[l], rank * 32, 32, DM, 0); break;
      case K_CQ: issue_rows(stage, 0, P.w_ca[l], rank * 32, 32, DM, 0); break;              // q part = first DM rows of in_proj
      case K_COUT: issue_rows(stage, 0, P.w_co[l], rank * 32, 32, DM, 0); break;
      case K_F1: issue_rows(stage, 0, P.w_f1[l], rank * FS + id.sub * 64, 64, DM, 0); break;
      case K_F2: issue_rows(stage, 0, P.w_f2[l], id.sub * 64, 64, FFN, rank * FS); break;    // K-split: columns of own hidden slice
      case K_HEAD: {
        const int r0 = rank * VSL, n = max(0, min(VSL, P.vocab - r0));
        issue_rows(stage, 0, P.w_out, r0, n, DM, 0);
        break;
      }
      case K_SELFK: case K_SELFV: {
        // keys 0..t-1 of head `rank` for images [sub*ips, ...): 64 B per key -> 4 chunks, padded rows of KVP elements
        const int which = id.kind == K_SELFV, nk = id.t;          // key t itself comes from shared memory
        const int g0 = id.sub * P.ips, gn = min(P.ips, G - g0);
        const int64_t plane = (int64_t)P.PT * DM;
        for (int c = threadIdx.x; c < gn * nk * 4; c += NT) {
          const int ch = c & 3, ku = (c >> 2) % nk, gi = (c >> 2) / nk;
          const int page = sm.pages[(g0 + gi) * 32 + ku / P.PT];
          const bf16* src = P.kv_pool + (((int64_t)page * P.layers + l) * 2 + which) * plane + (int64_t)(ku % P.PT) * DM + rank * HD + ch * 8;
          cp16(stage + ((size_t)(gi * P.maxT + ku) * KVP + ch * 8) * 2, src);
        }
        break;
      }
      case K_CROSSK: case K_CROSSV: {
        const int which = id.kind == K_CROSSV;
        const int g0 = id.sub * 2, gn = min(2, G - g0), S = P.S;
        const bf16* base = P.cross_kv + (int64_t)l * P.B * S * 2 * DM + which * DM + rank * HD;
        for (int c = threadIdx.x; c < gn * S * 4; c += NT) {
          const int ch = c & 3, ku = (c >> 2) % S, gi = (c >> 2) / S;
          cp16(stage + ((size_t)(gi * S + ku) * KVP + ch * 8) * 2, base + ((int64_t)(img0 + g0 + gi) * S + ku) * 2 * DM + ch * 8);
        }
        break;
      }
    }
  }
  cp_commit();     // always commit (possibly empty) so that wait_group accounting stays uniform
}

// one ring stage of a projection: out[g][col0 + 8w + ..] (+)= A[g][:] . Wstage[8w + n][:]   (warp w = n-tile w)
template <typename Store>
__device__ __forceinline__ void mma_stage(const uint8_t* stage, int n_tiles, const bf16* Ahi, const bf16* Alo, Store store) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= n_tiles) return;
  const bf16* W = reinterpret_cast<const bf16*>(stage) + (size_t)(warp * 8 + (lane >> 2)) * PITCH + 2 * (lane & 3);
  const bf16* Ah = Ahi + (size_t)(lane >> 2) * PITCH + 2 * (lane & 3);
  const bf16* Al = Alo + (size_t)(lane >> 2) * PITCH + 2 * (lane & 3);
  float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int k0 = 0; k0 < DM; k0 += 16) {
    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(W + k0), b1 = *reinterpret_cast<const uint32_t*>(W + k0 + 8);
    uint32_t ah[4], al[4];
    ah[0] = *reinterpret_cast<const uint32_t*>(Ah + k0); ah[1] = *reinterpret_cast<const uint32_t*>(Ah + 8 * PITCH + k0);
    ah[2] = *reinterpret_cast<const uint32_t*>(Ah + k0 + 8); ah[3] = *reinterpret_cast<const uint32_t*>(Ah + 8 * PITCH + k0 + 8);
    al[0] = *reinterpret_cast<const uint32_t*>(Al + k0); al[1] = *reinterpret_cast<const uint32_t*>(Al + 8 * PITCH + k0);
    al[2] = *reinterpret_cast<const uint32_t*>(Al + k0 + 8); al[3] = *reinterpret_cast<const uint32_t*>(Al + 8 * PITCH + k0 + 8);
    mma16816(c, ah, b0, b1);
    mma16816(c, al, b0, b1);
  }
  const int r0 = lane >> 2, cc = warp * 8 + 2 * (lane & 3);
  store(r0, cc, c[0]); store(r0, cc + 1, c[1]); store(r0 + 8, cc, c[2]); store(r0 + 8, cc + 1, c[3]);
}

// attention of one query (image g, own head) against nk keys held in a stage panel (padded bf16 rows) [+ one extra key in regs]
// scores -> sc[0..nk]; returns via out[32].  One warp.
__device__ __forceinline__ void attend_panel(const float* q /*smem 32, pre-scaled*/, const bf16* Kp, int nk, const uint8_t* padf,
                                             const float* k_extra, bool has_extra, float extra_bias, float* sc) {
  const int lane = threadIdx.x & 31;
  float qv[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) qv[j] = q[j];
  float mx = -INFINITY;
  for (int u = lane; u < nk; u += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(Kp + (size_t)u * KVP);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 raw = kr[c];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); s = fmaf(qv[c * 8 + 2 * i], f.x, s); s = fmaf(qv[c * 8 + 2 * i + 1], f.y, s); }
    }
    if (padf && padf[u]) s += 1.0f;
    sc[u] = s; mx = fmaxf(mx, s);
  }
  if (has_extra) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s = fmaf(qv[j], k_extra[j], s);
    s += extra_bias;
    if (lane == 0) sc[nk] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  __syncwarp();
  const int n = nk + (has_extra ? 1 : 0);
  float sum = 0.f;
  for (int u = lane; u < n; u += 32) { float e = expf(sc[u] - mx); sc[u] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) sc[n] = 1.0f / sum;     // normaliser parked behind the probabilities
  __syncwarp();
}

__device__ __forceinline__ float pv_panel(const bf16* Vp, int nk, const float* sc, const float* v_extra, bool has_extra) {
  const int lane = threadIdx.x & 31;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  int u = 0;
  for (; u + 4 <= nk; u += 4) {
    acc0 = fmaf(sc[u], __bfloat162float(Vp[(size_t)u * KVP + lane]), acc0);
    acc1 = fmaf(sc[u + 1], __bfloat162float(Vp[(size_t)(u + 1) * KVP + lane]), acc1);
    acc2 = fmaf(sc[u + 2], __bfloat162float(Vp[(size_t)(u + 2) * KVP + lane]), acc2);
    acc3 = fmaf(sc[u + 3], __bfloat162float(Vp[(size_t)(u + 3) * KVP + lane]), acc3);
  }
  for (; u < nk; ++u) acc0 = fmaf(sc[u], __bfloat162float(Vp[(size_t)u * KVP + lane]), acc0);
  float acc = (acc0 + acc1) + (acc2 + acc3);
  const int n = nk + (has_extra ? 1 : 0);
  if (has_extra) acc = fmaf(sc[nk], v_extra[lane], acc);
  return acc * sc[n];
}

__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(NT, 1) decode_cluster_kernel(const ClusterParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int img0 = cid * P.G;
  const int G = min(P.G, P.B - img0);          // images of this cluster (>= 1 by construction)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Smem sm = carve(smem_raw);
  const Sched sc(P.G, P.ips);
  const int sps = P.layers * sc.spl + 1;
  const int64_t total = (int64_t)(P.t_end - P.t_begin) * sps;
  const float scale = rsqrtf((float)HD);       // 1/sqrt(32)

  // one-time: zero the A operands (rows >= G stay zero), page ids, PAD flags of the already-known prefix
  for (int i = tid; i < 4 * GM * PITCH; i += NT) sm.a_hi[i] = __float2bfloat16_rn(0.f);   // a_hi,a_lo,f_hi,f_lo are contiguous
  for (int i = tid; i < GM * 32; i += NT) {
    const int g = i >> 5, j = i & 31;
    sm.pages[i] = (g < G && j < P.pages_per_seq) ? P.page_table[(int64_t)(img0 + g) * P.pages_per_seq + j] : 0;
  }
  for (int i = tid; i < GM * 256; i += NT) {
    const int g = i >> 8, u = i & 255;
    sm.padflag[i] = (g < G && u < P.t_begin) ? (P.tokens[(int64_t)(img0 + g) * P.tokens_ld + u] == P.pad_idx) : 0;
  }
  __syncthreads();

  int64_t cons = 0;                                     // next stage to consume
  for (int i = 0; i < NS - 1; ++i) issue_stage(P, sc, sm, i, total, rank, img0, G);
  // acquire(): stage `cons` has landed for every thread and slot (cons-1)%NS is free -> refill it
  auto acquire = [&]() -> const uint8_t* {
    cp_wait<NS - 2>();
    __syncthreads();
    issue_stage(P, sc, sm, cons + NS - 1, total, rank, img0, G);
    const uint8_t* st = sm.ring + (size_t)(cons % NS) * STAGE_BYTES;
    ++cons;
    return st;
  };
  // rows g = warp, warp+8 of a [GM][..] activation are owned by `warp` in the gather / LayerNorm phases
  auto ln_gather = [&](const float* lnw, const float* lnb) {
    // x = LN(xres + gathered yslice); writes xres and the hi/lo A operand.  lane l <-> column 32p + l of peer p
    for (int g = warp; g < G; g += 8) {
      float v[CS]; float s = 0.f;
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        const float* peer = cluster.map_shared_rank(sm.yslice, p);
        v[p] = sm.xres[g * DM + 32 * p + lane] + peer[g * 32 + lane];
        s += v[p];
      }
      const float mean = warp_sum(s) * (1.0f / DM);
      float q = 0.f;
#pragma unroll
      for (int p = 0; p < CS; ++p) { const float d = v[p] - mean; q += d * d; }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / DM) + 1e-5f);
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        const int c = 32 * p + lane;
        const float xn = (v[p] - mean) * rstd * __ldg(lnw + c) + __ldg(lnb + c);
        sm.xres[g * DM + c] = xn;
        split_store(sm.a_hi, sm.a_lo, g * PITCH + c, xn);
      }
    }
  };
  auto gather_o = [&]() {       // A operand <- concatenated head outputs
    for (int g = warp; g < G; g += 8) {
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        const float* peer = cluster.map_shared_rank(sm.oslice, p);
        split_store(sm.a_hi, sm.a_lo, g * PITCH + 32 * p + lane, peer[g * 32 + lane]);
      }
    }
  };

  for (int t = P.t_begin; t < P.t_end; ++t) {
    // ---- embedding + positional row (model.py:98-101); PAD flag of the token at position t --------------
    for (int g = warp; g < G; g += 8) {
      const int tok = __ldcg(P.tokens + (int64_t)(img0 + g) * P.tokens_ld + t);
      if (lane == 0) sm.padflag[g * 256 + t] = (tok == P.pad_idx);
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        const int c = 32 * p + lane;
        const float x = __ldg(P.emb + (int64_t)tok * DM + c) + __ldg(P.pos + (int64_t)t * DM + c);
        sm.xres[g * DM + c] = x;
        split_store(sm.a_hi, sm.a_lo, g * PITCH + c, x);
      }
    }
    for (int l = 0; l < P.layers; ++l) {
      // ---- self-attention in-proj: own head's q | k (stage A) and v (stage B) -----------------------------
      {
        const uint8_t* st = acquire();
        const float* bi = P.b_in[l];
        mma_stage(st, 8, sm.a_hi, sm.a_lo, [&](int r, int c, float v) {
          const int grow = (c < 32) ? rank * HD + c : DM + rank * HD + (c - 32);
          sm.qkv[r * 96 + c] = v + __ldg(bi + grow);
        });
        st = acquire();
        mma_stage(st, 4, sm.a_hi, sm.a_lo, [&](int r, int c, float v) { sm.qkv[r * 96 + 64 + c] = v + __ldg(bi + 2 * DM + rank * HD + c); });
      }
      __syncthreads();
      // append k_t, v_t (bf16) to the paged cache; keep the ROUNDED values for this step's own key
      for (int i = tid; i < G * 64; i += NT) {
        const int g = i >> 6, c = i & 63;           // c < 32: k, else v
        const bf16 h = __float2bfloat16_rn(sm.qkv[g * 96 + 32 + c]);
        sm.qkv[g * 96 + 32 + c] = __bfloat162float(h);
        const int page = sm.pages[g * 32 + t / P.PT];
        P.kv_pool[(((int64_t)page * P.layers + l) * 2 + (c >> 5)) * ((int64_t)P.PT * DM) + (int64_t)(t % P.PT) * DM + rank * HD + (c & 31)] = h;
      }
      // ---- self-attention, head `rank`: K panels then V panels (probabilities parked in sm.scores) --------
      {
        // q pre-scaled into sm.qc (reused as the query buffer)
        for (int i = tid; i < G * 32; i += NT) sm.qc[i] = sm.qkv[(i >> 5) * 96 + (i & 31)] * scale;
        for (int s = 0; s < sc.nS; ++s) {
          const uint8_t* st = acquire();            // the barrier inside also publishes sm.qc / rounded k,v
          const int g0 = s * P.ips, gn = min(P.ips, G - g0);       // ips <= 8: at most one image per warp and stage
          if (warp < gn) {
            const int g = g0 + warp;
            attend_panel(sm.qc + g * 32, reinterpret_cast<const bf16*>(st) + (size_t)warp * P.maxT * KVP, t, sm.padflag + g * 256,
                         sm.qkv + g * 96 + 32, true, sm.padflag[g * 256 + t] ? 1.0f : 0.0f, sm.scores + (size_t)warp * SCR);
          }
          st = acquire();                           // the matching V panel; the probabilities stay in this warp's score row
          if (warp < gn) {
            const int g = g0 + warp;
            sm.oslice[g * 32 + lane] = pv_panel(reinterpret_cast<const bf16*>(st) + (size_t)warp * P.maxT * KVP, t,
                                                sm.scores + (size_t)warp * SCR, sm.qkv + g * 96 + 64, true);
          }
        }
      }
      cluster.sync();                                                            // #1: head outputs visible
      gather_o();
      // ---- self out-proj slice -> yslice ---------------------------------------------------------------------
      {
        const uint8_t* st = acquire();
        const float* bo = P.b_so[l];
        mma_stage(st, 4, sm.a_hi, sm.a_lo, [&](int r, int c, float v) { sm.yslice[r * 32 + c] = v + __ldg(bo + rank * 32 + c); });
      }
      cluster.sync();                                                            // #2
      ln_gather(P.ln1w[l], P.ln1b[l]);
      // ---- cross-attention query slice -------------------------------------------------------------------------
      {
        const uint8_t* st = acquire();
        const float* bc = P.b_ca[l];
        mma_stage(st, 4, sm.a_hi, sm.a_lo, [&](int r, int c, float v) { sm.qc[r * 32 + c] = (v + __ldg(bc + rank * 32 + c)) * scale; });
      }
      // ---- cross-attention over the S memory keys (2 images per panel) ---------------------------------------
      for (int s = 0; s < sc.nC; ++s) {
        const uint8_t* st = acquire();
        const int gn = min(2, G - 2 * s);
        if (warp < gn) {
          const int g = 2 * s + warp;
          attend_panel(sm.qc + g * 32, reinterpret_cast<const bf16*>(st) + (size_t)warp * P.S * KVP, P.S, nullptr, nullptr, false, 0.f,
                       sm.scores + (size_t)warp * SCR);
        }
        st = acquire();
        if (warp < gn) {
          const int g = 2 * s + warp;
          sm.oslice[g * 32 + lane] = pv_panel(reinterpret_cast<const bf16*>(st) + (size_t)warp * P.S * KVP, P.S, sm.scores + (size_t)warp * SCR, nullptr, false);
        }
      }
      cluster.sync();                                                            // #3
      gather_o();
      {
        const uint8_t* st = acquire();
        const float* bo = P.b_co[l];
        mma_stage(st, 4, sm.a_hi, sm.a_lo, [&](int r, int c, float v) { sm.yslice[r * 32 + c] = v + __ldg(bo + rank * 32 + c); });
      }
      cluster.sync();                                                            // #4
      ln_gather(P.ln2w[l], P.ln2b[l]);
      // ---- FFN1: own 256 hidden units, ReLU, kept local as the FFN2 operand -----------------------------------
      for (int s = 0; s < 4; ++s) {
        const uint8_t* st = acquire();
        const float* b1 = P.b_f1[l];
        mma_stage(st, 8, sm.a_hi, sm.a_lo, [&](int r, int c, float v) {
          const int hcol = s * 64 + c;
          split_store(sm.f_hi, sm.f_lo, r * PITCH + hcol, fmaxf(v + __ldg(b1 + rank * FS + hcol), 0.f));
        });
      }
      // ---- FFN2 as a K-split: partial sums over the own hidden slice ------------------------------------------
      for (int s = 0; s < 4; ++s) {
        const uint8_t* st = acquire();
        mma_stage(st, 8, sm.f_hi, sm.f_lo, [&](int r, int c, float v) { sm.ypart[r * DM + s * 64 + c] = v; });
      }
      cluster.sync();                                                            // #5: partials visible
      for (int g = warp; g < G; g += 8) {                                       // reduce-scatter: own 32 columns
        float a = __ldg(P.b_f2[l] + rank * 32 + lane);
#pragma unroll
        for (int p = 0; p < CS; ++p) a += cluster.map_shared_rank(sm.ypart, p)[g * DM + rank * 32 + lane];
        sm.yslice[g * 32 + lane] = a;
      }
      cluster.sync();                                                            // #6
      ln_gather(P.ln3w[l], P.ln3b[l]);
    }
    // ---- vocabulary head: own 40 rows -> global step logits ------------------------------------------------------
    {
      const uint8_t* st = acquire();
      const int r0 = rank * VSL, nrows = max(0, min(VSL, P.vocab - r0));
      mma_stage(st, (nrows + 7) / 8, sm.a_hi, sm.a_lo, [&](int r, int c, float v) {
        if (r < G && c < nrows) {
          const float lg = v + __ldg(P.b_out + r0 + c);
          P.step_logits[(int64_t)(img0 + r) * P.vocab + r0 + c] = lg;
          if (P.logits_out) P.logits_out[(int64_t)(img0 + r) * P.logits_img_stride + (int64_t)(t + P.logits_row_offset) * P.vocab + r0 + c] = lg;
        }
      });
    }
    __threadfence();
    cluster.sync();                                                              // #7: all logits in L2
    // ---- select: CTA `rank` serves images rank, rank+8 (whole CTA, so the branch is uniform) ----------------
    for (int g = rank; g < G && !(P.forced && !(P.confs && (t % 4 == 0))); g += CS) {
      const int V = P.vocab;
      int Vp2 = 1; while (Vp2 < V) Vp2 <<= 1;
      float* lg = sm.scores; float* srt = sm.scores + V;       // V + Vp2 <= 8*SCR floats
      for (int i = tid; i < V; i += NT) lg[i] = __ldcg(P.step_logits + (int64_t)(img0 + g) * V + i);
      __syncthreads();
      const bool sample = (P.top_k != 0 || P.top_p != 1.0f) && P.uniforms != nullptr;
      const float u = sample ? P.uniforms[(int64_t)(img0 + g) * P.uniforms_ld + t] : 0.f;
      int token; float conf;
      select_from_logits(lg, srt, V, Vp2, P.top_k, P.top_p, sample, u, token, conf);
      if (tid == 0) {
        if (!P.forced) P.tokens[(int64_t)(img0 + g) * P.tokens_ld + t + 1] = token;
        if (P.confs && (t % 4 == 0)) P.confs[(int64_t)(img0 + g) * P.confs_ld + t / 4] = conf;
      }
      __syncthreads();
    }
    __threadfence();
    cluster.sync();                                                              // #8: tokens published
  }
  cp_wait<0>();
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------
int decode_cluster_supported(const mdc_model* m, const mdc_decode_state* st, int t_end) {
  const mdc_dims& d = m->d;
  if (d.precision != MDC_BF16 || d.dim != DM || d.dec_heads != CS || d.dec_ffn != FFN) return 0;
  if (d.dec_layers < 1 || d.dec_layers > 8 || d.vocab > CS * VSL || d.n_patches + 2 > SCR) return 0;
  if (st->x_override || st->pos_override) return 0;
  if (d.n_patches * KVP * 2 * 2 > STAGE_BYTES) return 0;          // two images' cross panels per stage
  if (t_end > 256 || st->pages_per_seq > 32) return 0;
  if (getenv("MDC_DECODE_BACKEND") && !strcmp(getenv("MDC_DECODE_BACKEND"), "generic")) return 0;
  return 1;
}

size_t decode_cluster_scratch_bytes(const mdc_model* m, int B) { return (size_t)B * m->d.vocab * sizeof(float); }

int decode_cluster_launch(mdc_model* m, const mdc_decode_state* st, int t_begin, int t_end, void* logits_scratch, cudaStream_t s) {
  mdc_ctx* ctx = m->ctx; const mdc_dims& d = m->d;
  const void** gw = m->w + MDC_ENC_GLOBAL_SLOTS + d.enc_depth * MDC_ENC_BLOCK_SLOTS;
  const void** lw0 = gw + MDC_DEC_GLOBAL_SLOTS;
  ClusterParams P{};
  for (int l = 0; l < d.dec_layers; ++l) {
    const void** lw = lw0 + l * MDC_DEC_LAYER_SLOTS;
    P.w_in[l] = (const bf16*)lw[MDC_SA_IN_W]; P.b_in[l] = (const float*)lw[MDC_SA_IN_B];
    P.w_so[l] = (const bf16*)lw[MDC_SA_OUT_W]; P.b_so[l] = (const float*)lw[MDC_SA_OUT_B];
    P.ln1w[l] = (const float*)lw[MDC_LN1_W]; P.ln1b[l] = (const float*)lw[MDC_LN1_B];
    P.w_ca[l] = (const bf16*)lw[MDC_CA_IN_W]; P.b_ca[l] = (const float*)lw[MDC_CA_IN_B];
    P.w_co[l] = (const bf16*)lw[MDC_CA_OUT_W]; P.b_co[l] = (const float*)lw[MDC_CA_OUT_B];
    P.ln2w[l] = (const float*)lw[MDC_LN2_W]; P.ln2b[l] = (const float*)lw[MDC_LN2_B];
    P.w_f1[l] = (const bf16*)lw[MDC_FF1_W]; P.b_f1[l] = (const float*)lw[MDC_FF1_B];
    P.w_f2[l] = (const bf16*)lw[MDC_FF2_W]; P.b_f2[l] = (const float*)lw[MDC_FF2_B];
    P.ln3w[l] = (const float*)lw[MDC_LN3_W]; P.ln3b[l] = (const float*)lw[MDC_LN3_B];
  }
  P.emb = (const float*)gw[MDC_EMB]; P.pos = (const float*)gw[MDC_DEC_POS];
  P.w_out = (const bf16*)gw[MDC_OUT_W]; P.b_out = (const float*)gw[MDC_OUT_B];
  P.layers = d.dec_layers; P.vocab = d.vocab; P.S = d.n_patches; P.pad_idx = d.pad_idx;
  P.B = st->B;
  P.tokens = st->tokens; P.tokens_ld = st->tokens_ld;
  P.kv_pool = (bf16*)st->kv_pool; P.page_table = st->page_table; P.pages_per_seq = st->pages_per_seq; P.PT = d.page_tokens;
  P.cross_kv = (const bf16*)st->cross_kv;
  P.step_logits = (float*)logits_scratch;
  P.logits_out = st->logits; P.logits_img_stride = (int64_t)st->logits_ld * d.vocab; P.logits_row_offset = st->logits_row_offset;
  P.confs = st->confs; P.confs_ld = st->confs_ld;
  P.uniforms = st->uniforms; P.uniforms_ld = st->uniforms_ld; P.top_k = st->top_k; P.top_p = st->top_p; P.forced = st->forced;
  P.t_begin = t_begin; P.t_end = t_end;
  P.maxT = ((t_end + 7) / 8) * 8;
  if (P.maxT < 8) P.maxT = 8;
  // images per cluster: as many clusters as the device can keep resident, at most GM images each
  static int max_clusters = 0;
  if (!max_clusters) {
    MDC_CUDA(cudaFuncSetAttribute(decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    cudaLaunchConfig_t q{}; q.gridDim = dim3(CS * 32); q.blockDim = dim3(NT); q.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CS; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    q.attrs = a; q.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, decode_cluster_kernel, &q);
    if (e != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = ctx->sm_count / CS / 2; if (n < 1) n = 1; }
    max_clusters = n;
  }
  int G = (P.B + max_clusters - 1) / max_clusters;
  if (G > GM) G = GM;
  if (G < 1) G = 1;
  P.G = G;
  P.ips = STAGE_BYTES / (P.maxT * KVP * 2);
  if (P.ips > 8) P.ips = 8;
  if (P.ips < 1) MDC_FAIL(-2, "decode_cluster: key capacity %d does not fit a stage", P.maxT);
  const int n_clusters = (P.B + G - 1) / G;
  decode_cluster_kernel<<<n_clusters * CS, NT, SMEM_BYTES, s>>>(P);
  MDC_LAUNCH_CHECK(ctx);
  return 0;
}
