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
  int pt_shift;                            // log2(PT)
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

// Incremental position of the PRODUCER in the static stage schedule (no div/mod on the hot path).
struct Cursor {
  int t, layer, j; bool head;
  __device__ __forceinline__ void advance(const Sched& sc, int layers) {
    if (head) { head = false; ++t; layer = 0; j = 0; }
    else if (++j == sc.spl) { j = 0; if (++layer == layers) head = true; }
  }
  __device__ __forceinline__ StageId id(const Sched& sc) const {
    StageId r; r.t = t; r.layer = layer; r.sub = 0;
    if (head) { r.kind = K_HEAD; return r; }
    int q = j;
    if (q < 2) { r.kind = q == 0 ? K_INA : K_INB; return r; } q -= 2;
    if (q < 2 * sc.nS) { r.kind = (q & 1) ? K_SELFV : K_SELFK; r.sub = q >> 1; return r; } q -= 2 * sc.nS;
    if (q == 0) { r.kind = K_SOUT; return r; } q -= 1;
    if (q == 0) { r.kind = K_CQ; return r; } q -= 1;
    if (q < 2 * sc.nC) { r.kind = (q & 1) ? K_CROSSV : K_CROSSK; r.sub = q >> 1; return r; } q -= 2 * sc.nC;
    if (q == 0) { r.kind = K_COUT; return r; } q -= 1;
    if (q < 4) { r.kind = K_F1; r.sub = q; return r; } q -= 4;
    r.kind = K_F2; r.sub = q; return r;
  }
};

// copy `nrows` weight rows (256 bf16 each, source pitch `spitch` elements) to stage rows [dst0, ...): 32 chunks of 16 B per row,
// one warp-pass = one row
__device__ __forceinline__ void issue_rows(uint8_t* stage, int dst0, const bf16* W, int nrows, int spitch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* src = W + (int64_t)warp * spitch + lane * 8;
  uint8_t* dst = stage + (size_t)(dst0 + warp) * (PITCH * 2) + lane * 16;
  for (int r = warp; r < nrows; r += 8) { cp16(dst, src); src += (int64_t)8 * spitch; dst += 8 * PITCH * 2; }
}

// copy nk keys (64 B = 4 chunks each) of one K or V panel: 4 lanes per key, division-free
__device__ __forceinline__ void issue_panel_contig(uint8_t* dst_panel, const bf16* src0, int64_t key_pitch, int nk) {
  const int ch = threadIdx.x & 3;
  for (int ku = threadIdx.x >> 2; ku < nk; ku += NT / 4)
    cp16(dst_panel + ((size_t)ku * KVP + ch * 8) * 2, src0 + (int64_t)ku * key_pitch + ch * 8);
}

__device__ void issue_stage(const ClusterParams& P, const Sched& sc, const Smem& sm, Cursor& cur, int slot, int rank, int img0, int G) {
  if (cur.t < P.t_end) {
    uint8_t* stage = sm.ring + (size_t)slot * STAGE_BYTES;
    const StageId id = cur.id(sc);
    const int l = id.layer;
    switch (id.kind) {
      case K_INA:
        issue_rows(stage, 0, P.w_in[l] + (int64_t)(rank * HD) * DM, 32, DM);               // q rows of head `rank`
        issue_rows(stage, 32, P.w_in[l] + (int64_t)(DM + rank * HD) * DM, 32, DM);          // k rows
        break;
      case K_INB: issue_rows(stage, 0, P.w_in[l] + (int64_t)(2 * DM + rank * HD) * DM, 32, DM); break;   // v rows
      case K_SOUT: issue_rows(stage, 0, P.w_so[l] + (int64_t)(rank * 32) * DM, 32, DM); break;
      case K_CQ: issue_rows(stage, 0, P.w_ca[l] + (int64_t)(rank * 32) * DM, 32, DM); break;             // q part = first DM rows of in_proj
      case K_COUT: issue_rows(stage, 0, P.w_co[l] + (int64_t)(rank * 32) * DM, 32, DM); break;
      case K_F1: issue_rows(stage, 0, P.w_f1[l] + (int64_t)(rank * FS + id.sub * 64) * DM, 64, DM); break;
      case K_F2: issue_rows(stage, 0, P.w_f2[l] + (int64_t)(id.sub * 64) * FFN + rank * FS, 64, FFN); break;   // K-split: own hidden columns
      case K_HEAD: {
        const int r0 = rank * VSL, n = max(0, min(VSL, P.vocab - r0));
        issue_rows(stage, 0, P.w_out + (int64_t)r0 * DM, n, DM);
        break;
      }
      case K_SELFK: case K_SELFV: {
        // keys 0..t-1 of head `rank` for images [sub*ips, ...); the step's own key t comes from shared memory
        const int which = id.kind == K_SELFV, nk = id.t;
        const int g0 = id.sub * P.ips, gn = min(P.ips, G - g0);
        const int64_t plane = (int64_t)P.PT * DM;
        const int ch = threadIdx.x & 3;
        for (int gi = 0; gi < gn; ++gi) {
          const int* pg = sm.pages + (g0 + gi) * 32;
          uint8_t* dstp = stage + (size_t)gi * P.maxT * KVP * 2;
          for (int ku = threadIdx.x >> 2; ku < nk; ku += NT / 4) {
            const bf16* src = P.kv_pool + (((int64_t)pg[ku >> P.pt_shift] * P.layers + l) * 2 + which) * plane +
                              (int64_t)(ku & (P.PT - 1)) * DM + rank * HD + ch * 8;
            cp16(dstp + ((size_t)ku * KVP + ch * 8) * 2, src);
          }
        }
        break;
      }
      case K_CROSSK: case K_CROSSV: {
        const int which = id.kind == K_CROSSV;
        const int g0 = id.sub * 2, gn = min(2, G - g0), S = P.S;
        const bf16* base = P.cross_kv + (int64_t)l * P.B * S * 2 * DM + which * DM + rank * HD;
        for (int gi = 0; gi < gn; ++gi)
          issue_panel_contig(stage + (size_t)gi * S * KVP * 2, base + (int64_t)(img0 + g0 + gi) * S * 2 * DM, 2 * DM, S);
        break;
      }
    }
    cur.advance(sc, P.layers);
  }
  cp_commit();     // always commit (possibly empty) so that wait_group accounting stays uniform
}

__device__ __forceinline__ void ldsm_x4(uint32_t* r, const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}

// One 16x8 output tile of a projection: C[g][n] = sum_k (Ahi+Alo)[g][k] * W[tile rows n][k], k = 0..255.
// Shared (NOT inlined) by every projection phase to keep the instruction footprint small.  Fragments come from
// ldmatrix (conflict-free with the 528-byte row pitch); four independent accumulator chains (hi/lo x even/odd
// k-step) hide the HMMA latency.
__device__ __noinline__ float4 mma_tile(const bf16* Wtile /*8 rows of the stage*/, const bf16* Ahi, const bf16* Alo) {
  const int lane = threadIdx.x & 31;
  const bf16* wp = Wtile + (size_t)(lane & 7) * PITCH + ((lane >> 3) & 1) * 8;
  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8, akof = (lane >> 4) * 8;
  const bf16* ah = Ahi + (size_t)arow * PITCH + akof;
  const bf16* al = Alo + (size_t)arow * PITCH + akof;
  float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
#pragma unroll 2
  for (int k0 = 0; k0 < DM; k0 += 32) {
    uint32_t b0, b1, b2, b3, h0[4], l0[4], h1[4], l1[4];
    ldsm_x2(b0, b1, wp + k0); ldsm_x2(b2, b3, wp + k0 + 16);
    ldsm_x4(h0, ah + k0); ldsm_x4(l0, al + k0); ldsm_x4(h1, ah + k0 + 16); ldsm_x4(l1, al + k0 + 16);
    mma16816(c0, h0, b0, b1); mma16816(c1, l0, b0, b1);
    mma16816(c2, h1, b2, b3); mma16816(c3, l1, b2, b3);
  }
  return make_float4((c0[0] + c2[0]) + (c1[0] + c3[0]), (c0[1] + c2[1]) + (c1[1] + c3[1]),
                     (c0[2] + c2[2]) + (c1[2] + c3[2]), (c0[3] + c2[3]) + (c1[3] + c3[3]));
}

// one ring stage of a projection: warp w computes n-tile w (8 output columns) and hands the 16x8 tile to `store(row, col, value)`
template <typename Store>
__device__ __forceinline__ void mma_stage(const uint8_t* stage, int n_tiles, const bf16* Ahi, const bf16* Alo, Store store) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= n_tiles) return;
  const float4 c = mma_tile(reinterpret_cast<const bf16*>(stage) + (size_t)warp * 8 * PITCH, Ahi, Alo);
  const int r0 = lane >> 2, cc = warp * 8 + 2 * (lane & 3);
  store(r0, cc, c.x); store(r0, cc + 1, c.y); store(r0 + 8, cc, c.z); store(r0 + 8, cc + 1, c.w);
}

// ---- attention of one query against a key panel, split over several warps (flash-decoding style) ------------------
// A warp owns keys [k_lo,k_hi) of one image (plus, for part 0 of self-attention, the step's own key held in shared
// memory).  Pass 1 parks exp(s - m_local) in the warp's score row and returns (m_local, l_local); pass 2 (after the V
// panel has landed) accumulates the un-normalised output for channel `lane`.  Partials are merged by attn_merge().
struct AttnJob {
  const float* q;            // smem, 32, pre-scaled
  const bf16* panel;         // K or V panel of this image (rows of KVP bf16)
  int k_lo, k_hi;
  const uint8_t* padf;       // PAD flags per key (self) or null (cross)
  const float* extra;        // own key / value (smem, 32 f32) or null
  float extra_bias;
  float* sc;                 // this warp's score row
};

__device__ __noinline__ float2 attn_scores(const AttnJob& j) {
  const int lane = threadIdx.x & 31;
  float qv[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) qv[i] = j.q[i];
  float mx = -INFINITY;
  for (int u = j.k_lo + lane; u < j.k_hi; u += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(j.panel + (size_t)u * KVP);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 raw = kr[c];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); s = fmaf(qv[c * 8 + 2 * i], f.x, s); s = fmaf(qv[c * 8 + 2 * i + 1], f.y, s); }
    }
    if (j.padf && j.padf[u]) s += 1.0f;
    j.sc[u - j.k_lo] = s; mx = fmaxf(mx, s);
  }
  const int n = j.k_hi - j.k_lo;
  if (j.extra) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s = fmaf(qv[i], j.extra[i], s);
    s += j.extra_bias;
    if (lane == 0) j.sc[n] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  __syncwarp();
  const int nn = n + (j.extra ? 1 : 0);
  float sum = 0.f;
  for (int u = lane; u < nn; u += 32) { float e = expf(j.sc[u] - mx); j.sc[u] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  return make_float2(mx, sum);
}

__device__ __noinline__ float attn_pv(const AttnJob& j) {
  const int lane = threadIdx.x & 31;
  const int n = j.k_hi - j.k_lo;
  const bf16* vp = j.panel + (size_t)j.k_lo * KVP + lane;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int u = 0;
  for (; u + 4 <= n; u += 4) {
    a0 = fmaf(j.sc[u], __bfloat162float(vp[(size_t)u * KVP]), a0);
    a1 = fmaf(j.sc[u + 1], __bfloat162float(vp[(size_t)(u + 1) * KVP]), a1);
    a2 = fmaf(j.sc[u + 2], __bfloat162float(vp[(size_t)(u + 2) * KVP]), a2);
    a3 = fmaf(j.sc[u + 3], __bfloat162float(vp[(size_t)(u + 3) * KVP]), a3);
  }
  for (; u < n; ++u) a0 = fmaf(j.sc[u], __bfloat162float(vp[(size_t)u * KVP]), a0);
  float acc = (a0 + a1) + (a2 + a3);
  if (j.extra) acc = fmaf(j.sc[n], j.extra[lane], acc);
  return acc;
}

// merge `parts` partial results of image g: part p at buf[(g*4+p)*36 + {0: m, 1: l, 2+c: o_c}]
__device__ __forceinline__ float attn_merge(const float* buf, int g, int parts) {
  const int lane = threadIdx.x & 31;
  float M = -INFINITY;
  for (int p = 0; p < parts; ++p) M = fmaxf(M, buf[(g * 4 + p) * 36]);
  float L = 0.f, o = 0.f;
  for (int p = 0; p < parts; ++p) {
    const float* b = buf + (g * 4 + p) * 36;
    if (b[1] > 0.f) { const float w = expf(b[0] - M); L = fmaf(w, b[1], L); o = fmaf(w, b[2 + lane], o); }
  }
  return o / L;
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

  int cons = 0;                                         // ring slot of the next stage to consume
  Cursor cur; cur.t = P.t_begin; cur.layer = 0; cur.j = 0; cur.head = false;
  for (int i = 0; i < NS - 1; ++i) issue_stage(P, sc, sm, cur, i, rank, img0, G);
  // acquire(): stage `cons` has landed for every thread and slot (cons-1)%NS is free -> refill it
  auto acquire = [&]() -> const uint8_t* {
    cp_wait<NS - 2>();
    __syncthreads();
    issue_stage(P, sc, sm, cur, (cons + NS - 1) & (NS - 1), rank, img0, G);
    const uint8_t* st = sm.ring + (size_t)cons * STAGE_BYTES;
    cons = (cons + 1) & (NS - 1);
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
        const int wpi = min(4, 8 / P.ips);          // warps per image inside a self-KV stage
        for (int s = 0; s < sc.nS; ++s) {
          const uint8_t* st = acquire();            // the barrier inside also publishes sm.qc / rounded k,v
          const int g0 = s * P.ips, gn = min(P.ips, G - g0);
          const int gi = warp / wpi, part = warp % wpi;
          const bool active = gi < gn;
          AttnJob job{};
          float2 ml = make_float2(-INFINITY, 0.f);
          if (active) {
            const int g = g0 + gi, chunk = (t + wpi - 1) / wpi;
            job.q = sm.qc + g * 32; job.panel = reinterpret_cast<const bf16*>(st) + (size_t)gi * P.maxT * KVP;
            job.k_lo = min(t, part * chunk); job.k_hi = min(t, (part + 1) * chunk);
            job.padf = sm.padflag + g * 256;
            job.extra = (part == 0) ? sm.qkv + g * 96 + 32 : nullptr;
            job.extra_bias = sm.padflag[g * 256 + t] ? 1.0f : 0.0f;
            job.sc = sm.scores + (size_t)warp * SCR;
            ml = attn_scores(job);
          }
          st = acquire();                           // the matching V panel; the probabilities stay in this warp's score row
          if (active) {
            const int g = g0 + gi;
            job.panel = reinterpret_cast<const bf16*>(st) + (size_t)gi * P.maxT * KVP;
            if (job.extra) job.extra = sm.qkv + g * 96 + 64;
            const float o = attn_pv(job);
            float* pb = sm.ypart + (g * 4 + part) * 36;        // ypart is idle between FFN reductions
            if (lane == 0) { pb[0] = ml.x; pb[1] = ml.y; }
            pb[2 + lane] = o;
          }
        }
        __syncthreads();
        for (int g = warp; g < G; g += 8) sm.oslice[g * 32 + lane] = attn_merge(sm.ypart, g, wpi);
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
        const int gi = warp >> 2, part = warp & 3;   // 4 warps per image, 2 images per stage
        const bool active = gi < gn;
        AttnJob job{};
        float2 ml = make_float2(-INFINITY, 0.f);
        if (active) {
          const int g = 2 * s + gi, chunk = (P.S + 3) / 4;
          job.q = sm.qc + g * 32; job.panel = reinterpret_cast<const bf16*>(st) + (size_t)gi * P.S * KVP;
          job.k_lo = min(P.S, part * chunk); job.k_hi = min(P.S, (part + 1) * chunk);
          job.sc = sm.scores + (size_t)warp * SCR;
          ml = attn_scores(job);
        }
        st = acquire();
        if (active) {
          const int g = 2 * s + gi;
          job.panel = reinterpret_cast<const bf16*>(st) + (size_t)gi * P.S * KVP;
          const float o = attn_pv(job);
          float* pb = sm.ypart + (g * 4 + part) * 36;
          if (lane == 0) { pb[0] = ml.x; pb[1] = ml.y; }
          pb[2 + lane] = o;
        }
      }
      __syncthreads();
      for (int g = warp; g < G; g += 8) sm.oslice[g * 32 + lane] = attn_merge(sm.ypart, g, 4);
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
  if (t_end > 256 || st->pages_per_seq > 32 || (d.page_tokens & (d.page_tokens - 1)) != 0) return 0;
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
  P.pt_shift = 0; while ((1 << P.pt_shift) < d.page_tokens) ++P.pt_shift;
  P.ips = STAGE_BYTES / (P.maxT * KVP * 2);
  if (P.ips > 8) P.ips = 8;
  if (P.ips < 1) MDC_FAIL(-2, "decode_cluster: key capacity %d does not fit a stage", P.maxT);
  const int n_clusters = (P.B + G - 1) / G;
  decode_cluster_kernel<<<n_clusters * CS, NT, SMEM_BYTES, s>>>(P);
  MDC_LAUNCH_CHECK(ctx);
  return 0;
}
