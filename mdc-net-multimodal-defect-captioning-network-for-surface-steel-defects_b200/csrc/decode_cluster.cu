// decode_cluster.cu -- the whole autoregressive decode loop as ONE persistent, cluster-cooperative kernel
// (north_star (c)+(d): paged bf16 self-KV, L2/HBM-resident cross-K/V, LayerNorm + residual + token select fused;
//  "no collective inside the decode loop" -- here not even a kernel boundary).
//
// Decomposition (model width 256, 8 heads x 32, FFN 2048 -- the configuration inference_p.py:126-129 builds):
//   * a thread-block CLUSTER of 8 CTAs owns a group of G <= 8 images (G <= 16 in the two-column-block instantiation the
//     batch pipeline uses) for the entire loop; clusters never talk to each other, so there is no grid-wide
//     synchronisation at all.  Further groups are processed back to back.
//   * inside a cluster CTA r owns attention head r and 1/8 of every projection: in-proj rows of head r (q,k,v),
//     out-proj / cross-q / cross-out rows [32r,32r+32), FFN1 hidden units [256r,256r+256), FFN2 as a K-split over
//     the same hidden units, vocabulary rows [40r,40r+40).
//   * everything that is read from HBM/L2 -- weights, the paged self-KV cache, the resident cross-K/V -- is fetched
//     by a dedicated PRODUCER WARP with TMA (cp.async.bulk.tensor, hardware 128B/64B swizzle) into a 5 x 32 KB (4 x 32 KB
//     with 16 images) shared-memory ring guarded by full/empty mbarriers; the producer walks the static stage schedule and runs
//     ahead of the 8 consumer warps across phase, layer and step boundaries.
//   * projections run on the tensor cores with the roles swapped: the WEIGHT rows are the MMA M dimension
//     (mma.sync m16n8k16, A fragments by ldmatrix from the swizzled TMA tile), the images are the N = 8
//     dimension (one or two column blocks per weight fragment), so no MMA lane is wasted on padding.  The decode-loop weights are IEEE fp16 (mdc_dims.dec_loop_dtype: same
//     bytes as bf16, 8x smaller rounding) and so are the projection operands (activations rounded to fp16: 11 significant bits).
//     Attention keeps bf16 K/V (the caches) with bf16 queries and probabilities.
//   * activations (a few KB) are exchanged through DISTRIBUTED SHARED MEMORY, push style: the producer of a slice
//     writes it into every peer with st.async (...mbarrier::complete_tx), the consumer waits on a local mbarrier
//     for the expected byte count.  No cluster-wide barrier inside the loop.
//   * LayerNorm+residual are fused into the all-gathers; the head's logits go straight to the CTA that owns the
//     image, which runs the greedy / top-k / top-p select and broadcasts the token.
// The generic kernels in decode.cu remain the path for the fp32 token-exact mode and other geometries;
// mdc_decode_steps picks this kernel when the geometry matches.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#define MDC_SEL_SYNC() asm volatile("bar.sync 1, 256;" ::: "memory")
#include "select.cuh"
using namespace mdcsel;

namespace {

constexpr int CS = 8;              // cluster size == heads
constexpr int DM = 256;            // model width
constexpr int HD = 32;             // head width
constexpr int FFN = 2048;
constexpr int FS = FFN / CS;       // hidden units per CTA (256)
constexpr int NCT = 256;           // consumer threads (8 warps)
constexpr int NT = NCT + 32;       // + producer warp
constexpr int XP = DM + 8;         // fp16 elements per padded activation row (528 B: conflict-free ldmatrix)
constexpr int STAGE_BYTES = 16384;   // one ring stage: a 32-row weight block (32 x 512 B), the K or V panel of one image (S <= 256 keys x 64 B) or packed self-KV pages
constexpr int VSL = 40;            // vocab rows per CTA; 8*40 = 320 >= V
constexpr int PSTR = 36;           // floats per attention partial: m, l, -, -, o[32] (o is 16-byte aligned)
constexpr int NPART = 9;           // attention partials per image: one per warp (interleaved key tiles) + the step's own key
constexpr int NS_MAX = 10;

// ---- shared memory map (bytes from the 1024-aligned base), per instantiation -------------------------------------
// NB = 8-image column blocks per cluster pass: NB = 1 (up to 8 images per cluster: lowest latency, the serial path) or NB = 2
// (up to 16 images: every weight fragment a warp loads feeds two MMA column blocks and every exchange carries twice the
// images -- less SM-time per image, used by the batch pipeline).  With 16 images the per-group buffers double, so the ring
// has 4 stages instead of 5, the FFN2 receive buffer shares its bytes with the attention partials and the select scratch
// with the q/k staging (disjoint phases, see the ordering notes at the uses).
// OVL: the overlays described above (always on with NB = 2; with NB = 1 they make the two-CTAs-per-SM layout fit).
template <int NB, int NSTG, bool OVL>
struct Lay {
  static constexpr int GMX = 8 * NB;                             // max images per cluster pass
  static constexpr int NS = NSTG;                                // ring stages (160 KB in flight at NB = 1: +1.2 % over 4, same-box A/B)
  static constexpr int ACT_BYTES = GMX * XP * 2;
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_XH = OFF_RING + NS * STAGE_BYTES;      // LN output (fp16 projection operand)
  static constexpr int OFF_OH = OFF_XH + ACT_BYTES;               // gathered attention output (written by peers)
  static constexpr int OFF_FH = OFF_OH + ACT_BYTES;               // own FFN hidden slice
  static constexpr int OFF_XRES = OFF_FH + ACT_BYTES;             // [GMX][DM] f32 residual stream
  static constexpr int OFF_YRECV = OFF_XRES + GMX * DM * 4;       // [GMX][DM] f32 all-gathered projection output (peers write)
  static constexpr int OFF_PART = OFF_YRECV + GMX * DM * 4;       // [GMX][NPART][PSTR] f32 attention partials
  static constexpr int PART_BYTES = GMX * NPART * PSTR * 4;
  static constexpr int F2_BYTES = CS * GMX * 32 * 4;              // [CS src][GMX][32] f32 FFN2 partial sums (peers write)
  static constexpr int OFF_F2RECV = !OVL ? OFF_PART + PART_BYTES : OFF_PART;
  static constexpr int OFF_QS = !OVL ? OFF_F2RECV + F2_BYTES : OFF_PART + (PART_BYTES > F2_BYTES ? PART_BYTES : F2_BYTES);   // [GMX][32] f32 scaled query of the own head
  static constexpr int OFF_KNEW = OFF_QS + GMX * 32 * 4;          // [GMX][32] f32 this step's key (bf16-rounded)
  static constexpr int OFF_VNEW = OFF_KNEW + GMX * 32 * 4;
  static constexpr int OFF_YTMP = OFF_VNEW + GMX * 32 * 4;        // [GMX][32] f32 own slice of a projection before the push
  static constexpr int OFF_QH = OFF_YTMP + GMX * 32 * 4;          // [GMX][32] bf16 scaled query (MMA A operand)
  static constexpr int OFF_LRECV = OFF_QH + GMX * 32 * 2;         // [NB][CS src][VSL] f32 logits of the images this CTA selects for
  static constexpr int SEL_BYTES = (CS * VSL + 512) * 4;          // select scratch: 320 + 512 floats
  static constexpr int OFF_SEL = !OVL ? OFF_LRECV + NB * CS * VSL * 4 : OFF_QS;
  static constexpr int OFF_TOK = !OVL ? OFF_SEL + SEL_BYTES : OFF_LRECV + NB * CS * VSL * 4;    // [2][GMX] int32 (double buffered by step parity)
  static constexpr int OFF_PAGES = OFF_TOK + 2 * GMX * 4;         // [GMX][32] int32
  static constexpr int OFF_PADF = OFF_PAGES + GMX * 32 * 4;       // [GMX][8] u32: bit u of an image's 256-bit row = token u is PAD
  static constexpr int OFF_HASPAD = OFF_PADF + GMX * 32;          // [GMX] int32: any PAD token among the keys so far
  static constexpr int OFF_BARS = OFF_HASPAD + GMX * 4;           // mbarriers
  static constexpr int NBARS = 2 * NS_MAX + 5;
  static constexpr int SMEM_USED = OFF_BARS + NBARS * 8;
  static constexpr int SMEM_BYTES = SMEM_USED + 1024;             // + alignment slack
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(OFF_XH % 16 == 0 && OFF_XRES % 16 == 0 && OFF_PART % 16 == 0 && OFF_QS % 16 == 0 && OFF_BARS % 8 == 0, "alignment");
  static_assert(!OVL || 4 * GMX * 32 * 4 >= SEL_BYTES, "select scratch must fit the q/k/v/y staging it shares");
};

enum { BAR_FULL = 0, BAR_EMPTY = NS_MAX, BAR_O = 2 * NS_MAX, BAR_Y, BAR_F2, BAR_LG, BAR_TOK, BAR_COUNT };

struct __align__(64) FusedParams {
  CUtensorMap m_in[8], m_so[8], m_ca[8], m_co[8], m_f1[8], m_f2[8];
  CUtensorMap m_head, m_head8, m_ckv, m_pool;
  const float* b_in[8]; const float* b_so[8]; const float* ln1w[8]; const float* ln1b[8];
  const float* b_ca[8]; const float* b_co[8]; const float* ln2w[8]; const float* ln2b[8];
  const float* b_f1[8]; const float* b_f2[8]; const float* ln3w[8]; const float* ln3b[8];
  const float* emb; const float* pos; const float* b_out;
  int layers, vocab, S, pad_idx;
  int B, G, n_groups;
  int32_t* tokens; int tokens_ld;
  bf16* kv_pool; const int32_t* page_table; int pages_per_seq, PT;
  float* logits_out; int64_t logits_img_stride; int logits_row_offset;
  float* confs; int confs_ld;
  const float* uniforms; int uniforms_ld; int top_k; float top_p; int forced;
  int t_begin, t_end;
  long long* trace; int trace_t;           // developer aid: phase timestamps of CTA 0 at step trace_t (MDC_DECODE_TRACE_PTR)
};

// ---- PTX helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU box.  try_wait carries a suspend-time
// hint: the waiting warp SLEEPS in hardware until the phase completes (or the hint expires) instead of spinning -- a spinning
// waiter (without the hint try_wait returns after a few hundred cycles: 1.2 G of the 6.5 G instructions of a 256-image launch were
// this loop, profiles/r2b) takes issue slots from the warps of the same SM sub-partition that have work, in particular from
// the second CTA that shares the SM in the two-CTAs-per-SM variant.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");      // hint: 1 ms
    if (done) break;
    if (++spins > 2000u) __trap();          // ~2 s
  }
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t v, uint32_t rbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 consumer warps
__device__ __forceinline__ void tma_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// projections: fp16 weights (A) x fp16 hi/lo activations (B), fp32 accumulate
__device__ __forceinline__ void mma16816_h(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}
// projection operands are fp16 (11 significant bits, 8x finer than bf16: tools/error_budget_cpu.py "activations fp16" moves the
// logit error by < 3e-4, where bf16 activations would cost 1.4e-2); |x| is clamped to the fp16 range first
__device__ __forceinline__ float clamp_h(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }
__device__ __forceinline__ void store_h(__half* dst, int idx, float x) { dst[idx] = __float2half_rn(clamp_h(x)); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h);
}

// One 16-row tile of a projection with the weights as the M operand:
//   acc[nb][0..3] = D[m0+g][img], D[m0+g][img+1], D[m0+g+8][img], D[m0+g+8][img+1], img = 8nb + 2q   (g = lane/4, q = lane%4)
// wblk: shared address of the TMA row block [4 k-blocks][R rows][128 B] (SWIZZLE_128B), fp16 weights; bh: fp16 activations
// [8 images][XP].  The fragment loads run two iterations ahead of the MMAs (the asm statements are volatile, so program order
// IS issue order: without the explicit skew every iteration pays the full ldmatrix latency; the buffer of iteration i is
// refilled for iteration i+2 as soon as its MMAs are issued); four accumulator chains.
template <int R, int NB>
__device__ __forceinline__ void mma_mtile(uint32_t wblk, int m0, uint32_t bh, float (&acc)[NB][4]) {
  const int lane = threadIdx.x & 31;
  const int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8, csel = lane >> 4, sw = lane & 7;
  const uint32_t a_base = wblk + row * 128;
  const uint32_t b_base = bh + (lane & 7) * (XP * 2) + (lane >> 3) * 16;
  float c[NB][4][4];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[nb][j][0] = c[nb][j][1] = c[nb][j][2] = c[nb][j][3] = 0.f; }
  uint32_t a0[2][4], a1[2][4], xb[2][NB][4];
#define MDC_MTILE_LOAD(I)                                                                                      \
  {                                                                                                            \
    constexpr int i_ = (I), kb_ = i_ >> 1, ch_ = (i_ & 1) * 4, buf_ = i_ & 1;                                  \
    ldsm_x4(a0[buf_], a_base + kb_ * (R * 128) + (((ch_ + csel) ^ sw) << 4));                                  \
    ldsm_x4(a1[buf_], a_base + kb_ * (R * 128) + (((ch_ + 2 + csel) ^ sw) << 4));                              \
    _Pragma("unroll") for (int nb_ = 0; nb_ < NB; ++nb_) ldsm_x4(xb[buf_][nb_], b_base + nb_ * (8 * XP * 2) + i_ * 64); \
  }
#define MDC_MTILE_MMA(I)                                                                                       \
  {                                                                                                            \
    constexpr int i_ = (I), buf_ = i_ & 1, par_ = (i_ & 1) * 2;                                                \
    _Pragma("unroll") for (int nb_ = 0; nb_ < NB; ++nb_) {                                                     \
      mma16816_h(c[nb_][par_], a0[buf_], xb[buf_][nb_][0], xb[buf_][nb_][1]);                                  \
      mma16816_h(c[nb_][par_ + 1], a1[buf_], xb[buf_][nb_][2], xb[buf_][nb_][3]);                              \
    }                                                                                                          \
  }
  MDC_MTILE_LOAD(0) MDC_MTILE_LOAD(1)
  MDC_MTILE_MMA(0) MDC_MTILE_LOAD(2)
  MDC_MTILE_MMA(1) MDC_MTILE_LOAD(3)
  MDC_MTILE_MMA(2) MDC_MTILE_LOAD(4)
  MDC_MTILE_MMA(3) MDC_MTILE_LOAD(5)
  MDC_MTILE_MMA(4) MDC_MTILE_LOAD(6)
  MDC_MTILE_MMA(5) MDC_MTILE_LOAD(7)
  MDC_MTILE_MMA(6) MDC_MTILE_MMA(7)
#undef MDC_MTILE_LOAD
#undef MDC_MTILE_MMA
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nb][j] = (c[nb][0][j] + c[nb][1][j]) + (c[nb][2][j] + c[nb][3][j]);
}

// ---- attention of ONE query per image on the tensor cores ------------------------------------------------------
// A key / value panel in shared memory holds key u at u*64 bytes, 16-byte chunk c stored at c ^ ((u >> 1) & 3)
// (TMA SWIZZLE_64B).  The query (bf16, like the K/V it meets -- the precision of the caches; tools/error_budget_cpu.py: "q bf16",
// "p bf16" move the logit error by < 5e-4) is row 0 of the M operand of S = q.K^T, so the score of keys 2q, 2q+1 of an 8-key
// tile lands in lanes (g = 0, q) -- exactly where the B fragment of P.V wants column 0: the softmax runs in registers, nothing
// goes through shared memory.  Lanes of the other row groups compute on zero rows; their values only reach output columns
// that are never read.  Scores are in log2 units (q is pre-scaled by log2(e)/sqrt(hd)), so p = exp2(s - m).
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// A fragments of q from its bf16 row (32 dims): aq[ks][h] covers dims 16*ks + 8*h + {2q, 2q+1} of MMA row 0 (other rows 0)
__device__ __forceinline__ void build_q_frag(const bf16* qh, uint32_t (&aq)[2][2]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, q4 = lane & 3;
  const bf16* src = qh + 2 * q4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(src + (i >> 1) * 16 + (i & 1) * 8);
    aq[i >> 1][i & 1] = g == 0 ? v : 0u;
  }
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One query against key tiles tl, tl + tstep (at most two, 16 keys each) of one (K, V) panel pair as ONE straight-line chunk: a
// tile past the end is clamped onto tile `tl` and masked, so there is no control flow between the fragment loads and the MMAs and
// all four K fragments and all four V fragments (which do not depend on the scores) are requested before the first MMA (volatile
// asm: program order is issue order) -- the chunk pays the ldmatrix latency once.  Requires tl < ntile.  On return m_out / l_out
// are valid in lanes 0-3; the un-normalised output of dims mt*16 + {g, g+8} sits in lanes with q == 0 as o[mt][0] and o[mt][2].
__device__ __forceinline__ void attn_chunk2(uint32_t kp, uint32_t vp, const uint32_t (&aq)[2][2], int tl, int tstep, int ntile, int nkeys,
                                            const uint32_t* padf, float& m_out, float& l_out, float (&o)[2][4]) {
  const int lane = threadIdx.x & 31, q4 = lane & 3;
  const uint32_t a_k0[4] = {aq[0][0], 0u, aq[0][1], 0u}, a_k1[4] = {aq[1][0], 0u, aq[1][1], 0u};
  const bool two = tl + tstep < ntile;
  const int k0a = tl * 16, k0b = (two ? tl + tstep : tl) * 16;
  uint32_t kb[2][2][4], va[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k0 = i ? k0b : k0a;
#pragma unroll
    for (int j = 0; j < 2; ++j) {              // 8-key n-tile j: one ldmatrix.x4 = the four dim chunks of keys k0+8j .. +7
      const int key = k0 + 8 * j + (lane & 7);
      ldsm_x4(kb[i][j], kp + key * 64 + (((lane >> 3) ^ ((key >> 1) & 3)) << 4));
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int key = (i ? k0b : k0a) + ((lane >> 4) & 1) * 8 + (lane & 7), sw = (key >> 1) & 3, dsel = (lane >> 3) & 1;
    ldsm_x4_trans(va[i][0], vp + key * 64 + ((dsel ^ sw) << 4));
    ldsm_x4_trans(va[i][1], vp + key * 64 + (((2 + dsel) ^ sw) << 4));
  }
  float sc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float c[4] = {0.f, 0.f, 0.f, 0.f}, d[4] = {0.f, 0.f, 0.f, 0.f};
      mma16816(c, a_k0, kb[i][j][0], kb[i][j][1]);
      mma16816(d, a_k1, kb[i][j][2], kb[i][j][3]);
      sc[i][2 * j] = c[0] + d[0];
      sc[i][2 * j + 1] = c[1] + d[1];
    }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int key = (i ? k0b : k0a) + (e >> 1) * 8 + 2 * q4 + (e & 1);
      if (padf != nullptr && ((padf[key >> 5] >> (key & 31)) & 1u)) sc[i][e] += LOG2E;   // float PAD-key bias +1.0 (Q7), in log2 units (padf: rare)
      if (key >= nkeys || (i == 1 && !two)) sc[i][e] = -INFINITY;        // tail of the last tile / clamped duplicate tile
    }
  float mx = fmaxf(fmaxf(fmaxf(sc[0][0], sc[0][1]), fmaxf(sc[0][2], sc[0][3])), fmaxf(fmaxf(sc[1][0], sc[1][1]), fmaxf(sc[1][2], sc[1][3])));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  // only row group g = 0 (lanes 0-3) holds scores -- key tl*16 is valid, so their mx is finite; the other lanes work on the all-zero
  // rows of the M operand (scores 0, own mx 0, p = 1: finite) and feed output columns that are never read, so no broadcast is needed
  float ls = 0.f;
  float o2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};          // second tile: its own accumulator chain
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { p[e] = ex2_approx(sc[i][e] - mx); ls += p[e]; }
    const uint32_t b0 = pack_bf16(p[0], p[1]), b1 = pack_bf16(p[2], p[3]);   // column 0 of the B operand lives in row group g = 0
    if (i == 0) { mma16816(o[0], va[0][0], b0, b1); mma16816(o[1], va[0][1], b0, b1); }
    else { mma16816(o2[0], va[1][0], b0, b1); mma16816(o2[1], va[1][1], b0, b1); }
  }
  ls += __shfl_xor_sync(0xffffffffu, ls, 1);
  ls += __shfl_xor_sync(0xffffffffu, ls, 2);
  l_out = ls;          // valid in lanes 0-3 (lane 0 stores the partial)
  m_out = mx;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[mt][e] += o2[mt][e];
}

// merge the partial results (m, l, -, -, o[32]) of one image; lane = channel.  `mask`: bit p set for every slot p (< NPART)
// written this phase; slots with l == 0 are empty.
__device__ __forceinline__ float attn_merge(const float* parts, uint32_t mask) {
  const int lane = threadIdx.x & 31;
  float m = -INFINITY, l = 0.f;
  if ((mask >> lane) & 1u) { m = parts[lane * PSTR]; l = parts[lane * PSTR + 1]; if (!(l > 0.f)) m = -INFINITY; }
  const float M = warp_max(m);
  const float w = (l > 0.f) ? exp2f(m - M) : 0.f;
  const float L = warp_sum(w * l);
  float o = 0.f;
#pragma unroll
  for (int p = 0; p < NPART; ++p) {
    const float wp = __shfl_sync(0xffffffffu, w, p);
    if (wp != 0.f) o = fmaf(wp, parts[p * PSTR + 4 + lane], o);      // warp-uniform
  }
  return o / L;
}

// Greedy select of one image by ONE warp, no block barrier: argmax (lowest index on ties, like torch.argmax) and the maximal
// softmax probability 1 / sum exp(l - max) (inference_p.py:77, 84-86 with top_k = 0, top_p = 1: the filter is the identity).
__device__ __forceinline__ void warp_greedy_select(const float* lg, int V, int& token, float& conf) {
  const int lane = threadIdx.x & 31;
  float bv = -INFINITY; int bi = 0x7fffffff;
  for (int i = lane; i < V; i += 32) { const float v = lg[i]; if (v > bv) { bv = v; bi = i; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  float sum = 0.f;
  for (int i = lane; i < V; i += 32) sum += expf(lg[i] - bv);
  token = bi; conf = 1.0f / warp_sum(sum);
}

// kTrace: developer build of the same kernel that stamps clock64() at phase boundaries (tools/decode_trace.py); the production
// instantiation carries no trace instructions.
// NSTG ring stages; MINB = 2: compact shared-memory layout and a register budget that let two CTAs (of two different clusters,
// i.e. two independent image groups) share an SM, so that one group's dependent phase chain fills the other's stalls.
template <bool kTrace, int NB, int NSTG, int MINB>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(NT, MINB) decode_fused_kernel(const __grid_constant__ FusedParams P) {
  using Y = Lay<NB, NSTG, NB == 2 || MINB == 2>;
  constexpr int GMX = Y::GMX, NS = Y::NS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cid = blockIdx.x / CS, n_clusters = gridDim.x / CS;

  __half* xh = (__half*)(smem + Y::OFF_XH); __half* fh = (__half*)(smem + Y::OFF_FH);      // projection operands (fp16)
  float* xres = (float*)(smem + Y::OFF_XRES); float* yrecv = (float*)(smem + Y::OFF_YRECV); float* f2recv = (float*)(smem + Y::OFF_F2RECV);
  float* qs = (float*)(smem + Y::OFF_QS); float* knew = (float*)(smem + Y::OFF_KNEW); float* vnew = (float*)(smem + Y::OFF_VNEW);
  float* ytmp = (float*)(smem + Y::OFF_YTMP); float* part = (float*)(smem + Y::OFF_PART);
  bf16* qh = (bf16*)(smem + Y::OFF_QH);
  float* lrecv = (float*)(smem + Y::OFF_LRECV); float* selbuf = (float*)(smem + Y::OFF_SEL);
  int* tokbuf = (int*)(smem + Y::OFF_TOK); int* pages = (int*)(smem + Y::OFF_PAGES); uint32_t* padbits = (uint32_t*)(smem + Y::OFF_PADF);
  int* haspad = (int*)(smem + Y::OFF_HASPAD);
  const uint32_t bars = sbase + Y::OFF_BARS;
  auto bar = [&](int i) -> uint32_t { return bars + i * 8; };

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bar(BAR_FULL + i), 1); mbar_init(bar(BAR_EMPTY + i), 8); }
    for (int i = BAR_O; i <= BAR_TOK; ++i) mbar_init(bar(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the ring and the activation operands once: rows >= G of the operands and the rows behind a partial key tile
  // of a K/V panel are multiplied by zeros and must be finite
  for (int i = tid; i < (NS * STAGE_BYTES + 3 * Y::ACT_BYTES) / 16; i += NT) reinterpret_cast<uint4*>(smem + Y::OFF_RING)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  cluster_sync_all();

  const float scale = rsqrtf((float)HD) * LOG2E;      // scores in log2 units: p = exp2(s - m)
  const int L = P.layers, S = P.S;
  const int nC = P.G;                                 // cross stages: K and V panel of ONE image each (by the full group size)
  // Self-KV stages are sized per step: with npg pages of keys an image is split into wpi = 1 / 2 / 4 / 8 chunks of two key tiles
  // (2 * wpi >= npg) and a 32 KB stage holds the K and V panels (wpi * 2 KB each) of ips = 8 / wpi images -- early steps pack 8 or 4
  // images into a stage.  A function of the step only (never of the batch: results stay batch-invariant).
  auto self_wpi = [](int npg) { return npg <= 2 ? 1 : (npg <= 4 ? 2 : (npg <= 8 ? 4 : 8)); };

  // ring cursors (producer and consumers keep their own copies)
  uint32_t slot = 0, phase = 0;
  uint32_t ph_o = 0, ph_y = 0, ph_f2 = 0, ph_lg = 0, ph_tok = 0;

  for (int grp = cid; grp < P.n_groups; grp += n_clusters) {
    const int img0 = grp * P.G;
    const int G = min(P.G, P.B - img0);
    // ---- per-group init: page ids, PAD flags of the known prefix ---------------------------------------------
    for (int i = tid; i < GMX * 32; i += NT) {
      const int g = i >> 5, j = i & 31;
      pages[i] = (g < G && j < P.pages_per_seq) ? P.page_table[(int64_t)(img0 + g) * P.pages_per_seq + j] : 0;
    }
    for (int i = tid; i < GMX * 8; i += NT) {      // PAD bits of the known prefix (tokens [0, t_begin))
      const int g = i >> 3, w = i & 7;
      uint32_t bits = 0;
      if (g < G)
        for (int b = 0; b < 32 && w * 32 + b < P.t_begin; ++b)
          if (P.tokens[(int64_t)(img0 + g) * P.tokens_ld + w * 32 + b] == P.pad_idx) bits |= 1u << b;
      padbits[i] = bits;
    }
    __syncthreads();
    if (tid < GMX) {
      uint32_t any = 0;
      for (int w = 0; w < 8; ++w) any |= padbits[tid * 8 + w];
      haspad[tid] = any != 0u;
    }
    __syncthreads();
    cluster_sync_all();          // no peer may push into this CTA before its buffers are set up for the group

    if (warp == 8) {
      // =================================== PRODUCER WARP ===================================================
      long long prod_wait = 0;
      auto acquire = [&]() -> uint32_t {           // wait until the consumers released the slot
        const long long c0 = kTrace ? clock64() : 0;
        mbar_wait(bar(BAR_EMPTY + slot), phase ^ 1);
        if (kTrace) prod_wait += clock64() - c0;
        return sbase + Y::OFF_RING + slot * STAGE_BYTES;
      };
      auto advance = [&]() { if (++slot == NS) { slot = 0; phase ^= 1; } };
      // one weight row block: 4 k-blocks of [R rows x 64 k]
      auto load_rows = [&](const CUtensorMap* m, uint32_t dst, uint32_t fb, int col0, int row0, int R) {
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_2d(m, fb, dst + kb * R * 128, col0 + kb * 64, row0);
      };
      for (int t = P.t_begin; t < P.t_end; ++t) {
        const int npg = (t + P.PT - 1) / P.PT;     // pages holding keys 0..t-1
        const int wpi_t = self_wpi(npg);
        const bool split_kv = wpi_t == 8;          // K panel and V panel of ONE image fill a stage each
        const int ips_t = split_kv ? 1 : 4 / wpi_t, nS = split_kv ? 2 * P.G : (P.G + ips_t - 1) / ips_t;
        const uint32_t self_panel = (uint32_t)wpi_t * 2048u;
        for (int l = 0; l < L; ++l) {
          for (int part = 0; part < 3; ++part) {   // in-proj: own head's q rows, k rows, v rows
            const uint32_t st = acquire(); const uint32_t fb = bar(BAR_FULL + slot);
            if (lane == 0) { mbar_expect_tx(fb, 32 * 512); load_rows(&P.m_in[l], st, fb, 0, part * DM + rank * HD, 32); }
            advance();
          }
          if (!split_kv) {
            for (int sg = 0; sg < nS; ++sg) {      // self-KV: K and V pages of images [sg*ips, ...), panels [gi][k|v]
              const int g0 = sg * ips_t, gn = max(0, min(ips_t, G - g0));
              const int ops = gn * 2 * npg;
              const uint32_t st = acquire(); const uint32_t fb = bar(BAR_FULL + slot);
              if (lane == 0) mbar_expect_tx(fb, ops * P.PT * 64);
              __syncwarp();
              for (int op = lane; op < ops; op += 32) {
                const int pn = op / npg, j = op - pn * npg;          // pn = gi*2 + which
                const int page = pages[(g0 + (pn >> 1)) * 32 + j];
                tma_2d(&P.m_pool, fb, st + pn * self_panel + j * (P.PT * 64), rank * HD, ((page * L + l) * 2 + (pn & 1)) * P.PT);
              }
              advance();
            }
          } else {
            for (int sg = 0; sg < nS; ++sg) {      // more than 8 pages of keys: the K pages of image sg/2 in one stage, its V pages in the next
              const int g = sg >> 1, which = sg & 1;
              const int ops = g < G ? npg : 0;
              const uint32_t st = acquire(); const uint32_t fb = bar(BAR_FULL + slot);
              if (lane == 0) mbar_expect_tx(fb, ops * P.PT * 64);
              __syncwarp();
              if (lane < ops) {
                const int page = pages[g * 32 + lane];
                tma_2d(&P.m_pool, fb, st + lane * (P.PT * 64), rank * HD, ((page * L + l) * 2 + which) * P.PT);
              }
              advance();
            }
          }
          {  // self out-proj rows, cross-q rows
            uint32_t st = acquire(); uint32_t fb = bar(BAR_FULL + slot);
            if (lane == 0) { mbar_expect_tx(fb, 32 * 512); load_rows(&P.m_so[l], st, fb, 0, rank * 32, 32); }
            advance();
            st = acquire(); fb = bar(BAR_FULL + slot);
            if (lane == 0) { mbar_expect_tx(fb, 32 * 512); load_rows(&P.m_ca[l], st, fb, 0, rank * 32, 32); }
            advance();
          }
          for (int sg = 0; sg < 2 * nC; ++sg) {    // cross K panel of image sg/2 in one stage, its V panel in the next
            const int g = sg >> 1, which = sg & 1;
            const uint32_t st = acquire(); const uint32_t fb = bar(BAR_FULL + slot);
            if (lane == 0) {
              if (g < G) {
                mbar_expect_tx(fb, S * 64);
                tma_2d(&P.m_ckv, fb, st, which * DM + rank * HD, (l * P.B + img0 + g) * S);
              } else mbar_expect_tx(fb, 0);
            }
            advance();
          }
          {  // cross out-proj rows
            const uint32_t st = acquire(); const uint32_t fb = bar(BAR_FULL + slot);
            if (lane == 0) { mbar_expect_tx(fb, 32 * 512); load_rows(&P.m_co[l], st, fb, 0, rank * 32, 32); }
            advance();
          }
          for (int s8 = 0; s8 < 8; ++s8) {         // FFN1: own hidden rows, 32 per stage
            const uint32_t st = acquire(); const uint32_t fb = bar(BAR_FULL + slot);
            if (lane == 0) { mbar_expect_tx(fb, 32 * 512); load_rows(&P.m_f1[l], st, fb, 0, rank * FS + s8 * 32, 32); }
            advance();
          }
          for (int s8 = 0; s8 < 8; ++s8) {         // FFN2: K-split over the own hidden columns, 32 output features per stage
            const uint32_t st = acquire(); const uint32_t fb = bar(BAR_FULL + slot);
            if (lane == 0) { mbar_expect_tx(fb, 32 * 512); load_rows(&P.m_f2[l], st, fb, rank * FS, s8 * 32, 32); }
            advance();
          }
        }
        {  // vocabulary head rows: [0,32) of the own 40 in one stage, [32,40) in the next
          uint32_t st = acquire(); uint32_t fb = bar(BAR_FULL + slot);
          if (lane == 0) { mbar_expect_tx(fb, 32 * 512); load_rows(&P.m_head, st, fb, 0, rank * VSL, 32); }
          advance();
          st = acquire(); fb = bar(BAR_FULL + slot);
          if (lane == 0) {
            mbar_expect_tx(fb, 8 * 512);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) tma_2d(&P.m_head8, fb, st + kb * 8 * 128, kb * 64, rank * VSL + 32);
          }
          advance();
        }
      }
      if (kTrace && lane == 0 && blockIdx.x == 0) P.trace[202] = prod_wait;
    } else {
      // =================================== CONSUMER WARPS ==================================================
      long long cons_wait = 0, exch_wait = 0;
      auto stage_wait = [&]() -> uint32_t {
        const long long c0 = kTrace ? clock64() : 0;
        mbar_wait(bar(BAR_FULL + slot), phase);
        if (kTrace) cons_wait += clock64() - c0;
        return sbase + Y::OFF_RING + slot * STAGE_BYTES;
      };
      auto stage_wait_next = [&]() -> uint32_t {     // the stage after the cursor (two-stage jobs: K panel, then V panel)
        const uint32_t s1 = slot + 1 == NS ? 0u : slot + 1, p1 = slot + 1 == NS ? phase ^ 1u : phase;
        const long long c0 = kTrace ? clock64() : 0;
        mbar_wait(bar(BAR_FULL + s1), p1);
        if (kTrace) cons_wait += clock64() - c0;
        return sbase + Y::OFF_RING + s1 * STAGE_BYTES;
      };
      auto xwait = [&](int which, uint32_t& ph) {    // wait for a push-style exchange
        const long long c0 = kTrace ? clock64() : 0;
        mbar_wait(bar(which), ph); ph ^= 1;
        if (kTrace) exch_wait += clock64() - c0;
      };
      // A stage is released by the warps that read it: the empty barrier counts 8 arrivals, a stage with k user warps gets
      // 8/k from each of them (`weight`); the other warps neither wait for the stage nor arrive, they only advance the cursor.
      auto stage_release = [&](uint32_t weight = 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive_n(bar(BAR_EMPTY + slot), weight);
        if (++slot == NS) { slot = 0; phase ^= 1; }
      };
      auto stage_skip = [&]() { if (++slot == NS) { slot = 0; phase ^= 1; } };
      const int fg = lane >> 2, fq = lane & 3;       // MMA fragment coordinates: row group, image pair
      int trace_n = 0;
      auto TRACE = [&](int t_now) { if (kTrace && tid == 0 && blockIdx.x == 0 && t_now == P.trace_t) P.trace[trace_n++] = clock64(); };
      auto FINE = [&](int l_now, int t_now, int idx) {   // trace build: fine stamps of layer 2 into trace[100 + idx] by lane 0 of the calling warp
        if (kTrace && lane == 0 && blockIdx.x == 0 && t_now == P.trace_t && l_now == 2) P.trace[100 + idx] = clock64();
      };
      const int gi_t = tid >> 5, c_t = tid & 31;     // (image [+ 8], channel) coordinates of the elementwise phases
      constexpr int APP0 = NCT - GMX * 8;           // first thread of the KV-append crew (the last GMX * 8 consumer threads)
      // push this CTA's [G][32] slice in ytmp into every peer's yrecv columns [32*rank, +32)
      auto push_y = [&](int img, float v) {
        if (img < G) {
          const uint32_t off = sbase + Y::OFF_YRECV + (img * DM + 32 * rank + c_t) * 4;
#pragma unroll
          for (int p = 0; p < CS; ++p) st_async_b32(mapa(off, p), __float_as_uint(v), mapa(bar(BAR_Y), p));
        }
      };
      auto wait_y = [&]() {
        if (tid == 0) mbar_expect_tx(bar(BAR_Y), G * DM * 4);
        xwait(BAR_Y, ph_y);
      };
      // x = LN(xres + yrecv): warp w owns images w (and w + 8); writes xres and the fp16 operand
      auto layer_norm = [&](const float* lnw, const float* lnb, bool fence_appends = false) {
        // this layer's KV append: the generic->async proxy fence (~1200 cycles) of the appending threads overlaps the exchange latency
        if (fence_appends && tid >= APP0 && tid - APP0 < G * 8) asm volatile("fence.proxy.async;" ::: "memory");
        float gw[8], gb[8];
        if (warp < G) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { gw[j] = __ldg(lnw + lane + 32 * j); gb[j] = __ldg(lnb + lane + 32 * j); }
        }
        wait_y();
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const int img = warp + 8 * nb;
          if (img < G) {
            float v[8]; float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[j] = xres[img * DM + lane + 32 * j] + yrecv[img * DM + lane + 32 * j]; s += v[j]; }
            const float mean = warp_sum(s) * (1.0f / DM);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; q += d * d; }
            const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / DM) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = lane + 32 * j;
              const float xn = (v[j] - mean) * rstd * gw[j] + gb[j];
              xres[img * DM + c] = xn;
              store_h(xh, img * XP + c, xn);
            }
          }
        }
        cbar();
      };
      // attention output of image `img` (lane = channel) -> fp16 pairs into every peer's oh columns [32*rank, +32); lane j < 16
      // sends channels 2j, 2j+1 to peers 0..3, lane 16 + j the same pair to peers 4..7
      auto push_o = [&](int img, float o) {
        const int j = lane & 15;
        const float v0 = __shfl_sync(0xffffffffu, o, 2 * j), v1 = __shfl_sync(0xffffffffu, o, 2 * j + 1);
        if (img < G) {
          const uint32_t val = pack_h2(clamp_h(v0), clamp_h(v1));
          const uint32_t off = sbase + Y::OFF_OH + (img * XP + 32 * rank + 2 * j) * 2;
          const int p0 = (lane >> 4) * 4;
#pragma unroll
          for (int p = 0; p < 4; ++p) st_async_b32(mapa(off, p0 + p), val, mapa(bar(BAR_O), p0 + p));
        }
      };
      auto wait_o = [&]() {
        if (tid == 0) mbar_expect_tx(bar(BAR_O), G * DM * 2);
        xwait(BAR_O, ph_o);
      };
      // a 32-row projection of the gathered operand bh -> ytmp (+bias) -> pushed to all peers
      auto proj32_push = [&](uint32_t bh, const float* bias) {
        float b0 = 0.f, b1 = 0.f;
        if (warp < 2) { b0 = __ldg(bias + rank * 32 + warp * 16 + fg); b1 = __ldg(bias + rank * 32 + warp * 16 + fg + 8); }
        if (warp < 2) {
          const uint32_t st = stage_wait();
          float acc[NB][4];
          mma_mtile<32, NB>(st, warp * 16, bh, acc);
          const int f = warp * 16 + fg;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            float* y = ytmp + (8 * nb + 2 * fq) * 32 + f;
            y[0] = acc[nb][0] + b0; y[32] = acc[nb][1] + b0; y[8] = acc[nb][2] + b1; y[40] = acc[nb][3] + b1;
          }
          stage_release(4);
        } else {
          stage_skip();
        }
        cbar();
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) push_y(gi_t + 8 * nb, ytmp[(gi_t + 8 * nb) * 32 + c_t]);
      };

      // one warp's share of one (image, head) attention job: key tiles tl, tl + tstep of the panels kp / vp -> partial slot `pslot`
      auto attend = [&](uint32_t kp, uint32_t vp, int g, int pslot, int tl, int tstep, int ntile, int nkeys, const uint32_t* padw) {
        float* pb = part + (g * NPART + pslot) * PSTR;
        if (tl < ntile) {
          uint32_t aq[2][2];
          build_q_frag(qh + g * 32, aq);
          float m_run, l_run;
          float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
          attn_chunk2(kp, vp, aq, tl, tstep, ntile, nkeys, padw, m_run, l_run, o);
          if (lane == 0) { pb[0] = m_run; pb[1] = l_run; }
          if ((lane & 3) == 0) {
            const int g8 = lane >> 2;
            pb[4 + g8] = o[0][0]; pb[4 + g8 + 8] = o[0][2];
            pb[4 + g8 + 16] = o[1][0]; pb[4 + g8 + 24] = o[1][2];
          }
        } else if (lane == 0) pb[1] = 0.f;
      };

      for (int t = P.t_begin; t < P.t_end; ++t) {
        const int wpi_t = self_wpi((t + P.PT - 1) / P.PT);
        const bool split_kv = wpi_t == 8;
        const int ips_t = split_kv ? 1 : 4 / wpi_t, nS = split_kv ? 2 * P.G : (P.G + ips_t - 1) / ips_t;
        const uint32_t self_panel = (uint32_t)wpi_t * 2048u;
        // ---- embedding + positional row (model.py:98-101); PAD flag of the token at position t --------------
        if (warp < G) {
          float pz[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pz[j] = __ldg(P.pos + (int64_t)t * DM + lane + 32 * j);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int img = warp + 8 * nb;
            if (img < G) {
              const int tok = (P.forced || t == P.t_begin) ? __ldcg(P.tokens + (int64_t)(img0 + img) * P.tokens_ld + t) : tokbuf[(t & 1) * GMX + img];
              if (lane == 0 && tok == P.pad_idx) { padbits[img * 8 + (t >> 5)] |= 1u << (t & 31); haspad[img] = 1; }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                const float x = __ldg(P.emb + (int64_t)tok * DM + c) + pz[j];
                xres[img * DM + c] = x;
                store_h(xh, img * XP + c, x);
              }
            }
          }
        }
        cbar();
        for (int l = 0; l < L; ++l) {
          TRACE(t);   // 0: layer start
          // ---- self-attention in-proj: own head's q (warps 0-1), k (warps 2-3), v (warps 4-5), one 32-row stage each ------
          {
            const float* bi = P.b_in[l];
            const int part = warp >> 1;                  // 0 q, 1 k, 2 v (3: no projection work)
            float b0 = 0.f, b1 = 0.f;
            if (part < 3) { const int r0 = part * DM + rank * HD + (warp & 1) * 16 + fg; b0 = __ldg(bi + r0); b1 = __ldg(bi + r0 + 8); }
#pragma unroll
            for (int pt = 0; pt < 3; ++pt) {
              if (part == pt) {
                const uint32_t st = stage_wait();
                float acc[NB][4];
                mma_mtile<32, NB>(st, (warp & 1) * 16, sbase + Y::OFF_XH, acc);
                const int f = (warp & 1) * 16 + fg;
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                  const int o = (8 * nb + 2 * fq) * 32 + f;
                  if (pt == 0) {       // q, pre-scaled
                    const float q0 = (acc[nb][0] + b0) * scale, q1 = (acc[nb][1] + b0) * scale, q2 = (acc[nb][2] + b1) * scale, q3 = (acc[nb][3] + b1) * scale;
                    const bf16 r0 = __float2bfloat16_rn(q0), r1 = __float2bfloat16_rn(q1), r2 = __float2bfloat16_rn(q2), r3 = __float2bfloat16_rn(q3);
                    qh[o] = r0; qh[o + 32] = r1; qh[o + 8] = r2; qh[o + 40] = r3;
                    // the step's own key meets the same bf16 query as the cached keys
                    qs[o] = __bfloat162float(r0); qs[o + 32] = __bfloat162float(r1); qs[o + 8] = __bfloat162float(r2); qs[o + 40] = __bfloat162float(r3);
                  } else {             // k / v, rounded to the cache precision
                    float* dst = pt == 1 ? knew : vnew;
                    dst[o] = __bfloat162float(__float2bfloat16_rn(acc[nb][0] + b0));
                    dst[o + 32] = __bfloat162float(__float2bfloat16_rn(acc[nb][1] + b0));
                    dst[o + 8] = __bfloat162float(__float2bfloat16_rn(acc[nb][2] + b1));
                    dst[o + 40] = __bfloat162float(__float2bfloat16_rn(acc[nb][3] + b1));
                  }
                }
                stage_release(4);
              } else stage_skip();
            }
          }
          cbar();
          TRACE(t);   // 1: in-proj done
          // append k_t, v_t (bf16) to the paged cache: 16-byte stores, 4 per (image, k|v), by the last warps.  The proxy fence that
          // orders them before the TMA reads of the page (one step later) costs ~1200 cycles; the same threads issue it while they
          // wait for the out-projection exchange (layer_norm), where its latency hides behind the exchange's.
          if (tid >= APP0 && tid - APP0 < G * 8) {
            const int at = tid - APP0;
            const int g = at >> 3, which = (at >> 2) & 1, ch = at & 3;
            const float* src = (which ? vnew : knew) + g * 32 + ch * 8;
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            uint4 o; o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
            const int page = pages[g * 32 + t / P.PT];
            bf16* dst = P.kv_pool + (((int64_t)page * L + l) * 2 + which) * ((int64_t)P.PT * DM) + (int64_t)(t % P.PT) * DM + rank * HD + ch * 8;
            *reinterpret_cast<uint4*>(dst) = o;
          }
          TRACE(t);   // 2: append
          // ---- self-attention, head `rank`: keys [0,t) from the paged cache.  An image's key tiles are interleaved over wpi_t warps (two
          //      tiles per warp at most: wpi_t * 2 * 16 >= t; a function of the step only -- never of the batch).  Up to 8 pages of keys:
          //      a 16 KB stage holds the K and V panels of 4 / wpi_t images and belongs to one half of the warps (even stages: warps
          //      0-3, odd stages: warps 4-7 -- two stages are worked on at a time); beyond that all 8 warps share one image, its K
          //      panel in one stage and its V panel in the next.  The step's own key (still in shared memory) is partial #8 --------
          {
            const int ntile = (t + 15) >> 4;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              const int img = warp + 8 * nb;
              if (img < G) {
                float s_own = warp_sum(qs[img * 32 + lane] * knew[img * 32 + lane]);
                if ((padbits[img * 8 + (t >> 5)] >> (t & 31)) & 1u) s_own += LOG2E;
                float* pb = part + (img * NPART + 8) * PSTR;
                if (lane == 0) { pb[0] = s_own; pb[1] = 1.0f; }
                pb[4 + lane] = vnew[img * 32 + lane];
              }
            }
            if (!split_kv) {
              const int wpi = wpi_t, half = warp >> 2, w4 = warp & 3, gi = w4 / wpi, tl = w4 - gi * wpi;
              for (int sg = 0; sg < nS; ++sg) {
                if ((sg & 1) != half) { stage_skip(); continue; }
                const int g = sg * ips_t + gi;
                const uint32_t st = stage_wait();
                if (g < G) {
                  const uint32_t kp = st + gi * 2 * self_panel;
                  attend(kp, kp + self_panel, g, tl, tl, wpi, ntile, t, haspad[g] ? padbits + g * 8 : nullptr);
                }
                stage_release(2);
              }
            } else {
              for (int g = 0; g < P.G; ++g) {
                const uint32_t kp = stage_wait(), vp = stage_wait_next();
                if (g < G) attend(kp, vp, g, warp, warp, 8, ntile, t, haspad[g] ? padbits + g * 8 : nullptr);
                stage_release(); stage_release();
              }
            }
          }
          cbar();
          TRACE(t);   // 3: self attention
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int img = warp + 8 * nb;
            push_o(img, img < G ? attn_merge(part + img * NPART * PSTR, ((1u << wpi_t) - 1u) | (1u << 8)) : 0.f);
          }
          wait_o();
          TRACE(t);   // 4: o gathered
          // ---- self out-proj slice -> all-gather -> LN1 ------------------------------------------------------------
          proj32_push(sbase + Y::OFF_OH, P.b_so[l]);
          TRACE(t);   // 5: out-proj pushed
          layer_norm(P.ln1w[l], P.ln1b[l], true);
          TRACE(t);   // 6: LN1

          // ---- cross-attention query slice -------------------------------------------------------------------------
          {
            const float* bc = P.b_ca[l];
            float b0 = 0.f, b1 = 0.f;
            if (warp < 2) { b0 = __ldg(bc + rank * 32 + warp * 16 + fg); b1 = __ldg(bc + rank * 32 + warp * 16 + fg + 8); }
            if (warp < 2) {
              const uint32_t st = stage_wait();
              float acc[NB][4];
              mma_mtile<32, NB>(st, warp * 16, sbase + Y::OFF_XH, acc);
              const int f = warp * 16 + fg;
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                bf16* q = qh + (8 * nb + 2 * fq) * 32 + f;
                q[0] = __float2bfloat16_rn((acc[nb][0] + b0) * scale); q[32] = __float2bfloat16_rn((acc[nb][1] + b0) * scale);
                q[8] = __float2bfloat16_rn((acc[nb][2] + b1) * scale); q[40] = __float2bfloat16_rn((acc[nb][3] + b1) * scale);
              }
              stage_release(4);
            } else stage_skip();
          }
          cbar();
          TRACE(t);   // 7: cross q
          // ---- cross-attention over the S memory keys: per image its K panel in one stage and its V panel in the next, key tiles
          //      interleaved over the 8 warps ------------------------------------------------------------------------------------
          {
            const int ntile = (S + 15) >> 4;
            for (int g = 0; g < nC; ++g) {
              if (warp == 0) FINE(l, t, 20 + g * 3);
              const uint32_t kp = stage_wait(), vp = stage_wait_next();
              if (warp == 0) FINE(l, t, 21 + g * 3);
              if (g < G) attend(kp, vp, g, warp, warp, 8, ntile, S, nullptr);
              stage_release(); stage_release();
              if (warp == 0) FINE(l, t, 22 + g * 3);
            }
          }
          cbar();
          TRACE(t);   // 8: cross attention partials
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int img = warp + 8 * nb;
            push_o(img, img < G ? attn_merge(part + img * NPART * PSTR, 0xffu) : 0.f);
          }
          wait_o();
          TRACE(t);   // 9: o gathered
          proj32_push(sbase + Y::OFF_OH, P.b_co[l]);
          TRACE(t);   // 10: cross out-proj pushed
          layer_norm(P.ln2w[l], P.ln2b[l]);
          TRACE(t);   // 11: LN2

          // ---- FFN1: own 256 hidden units in 8 stages of 32 rows (stage s -> warps 2(s%4), 2(s%4)+1: four stages are worked on at a
          //      time), ReLU, kept local as the FFN2 operand -------------------------------------------------------------------
          for (int s8 = 0; s8 < 8; ++s8) {
            const bool mine = (warp >> 1) == (s8 & 3);
            const int mt = warp & 1;
            if (mine) {
              const int h0 = rank * FS + s8 * 32 + mt * 16 + fg;
              const float b0 = __ldg(P.b_f1[l] + h0), b1 = __ldg(P.b_f1[l] + h0 + 8);
              if (warp == 0) FINE(l, t, (s8 >> 2) * 4 + 0);
              const uint32_t st = stage_wait();
              if (warp == 0) FINE(l, t, (s8 >> 2) * 4 + 1);
              float acc[NB][4];
              mma_mtile<32, NB>(st, mt * 16, sbase + Y::OFF_XH, acc);
              if (warp == 0) FINE(l, t, (s8 >> 2) * 4 + 2);
              const int h = s8 * 32 + mt * 16 + fg;
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                const int r = 8 * nb + 2 * fq;
                store_h(fh, r * XP + h, fmaxf(acc[nb][0] + b0, 0.f)); store_h(fh, (r + 1) * XP + h, fmaxf(acc[nb][1] + b0, 0.f));
                store_h(fh, r * XP + h + 8, fmaxf(acc[nb][2] + b1, 0.f)); store_h(fh, (r + 1) * XP + h + 8, fmaxf(acc[nb][3] + b1, 0.f));
              }
              stage_release(4);
              if (warp == 0) FINE(l, t, (s8 >> 2) * 4 + 3);
            } else stage_skip();
          }
          cbar();
          TRACE(t);   // 12: FFN1
          // ---- FFN2 as a K-split, 32 output features per stage: partial sums pushed straight to the CTA that owns the columns ----
          for (int s8 = 0; s8 < 8; ++s8) {
            const bool mine = (warp >> 1) == (s8 & 3);
            const int mt = warp & 1;
            if (mine) {
              const uint32_t st = stage_wait();
              float acc[NB][4];
              mma_mtile<32, NB>(st, mt * 16, sbase + Y::OFF_FH, acc);
              const int feat = s8 * 32 + mt * 16 + fg;                 // output feature of acc[.][0..1]; +8 for acc[.][2..3]
              const uint32_t peer = (uint32_t)s8;                       // stage s8 = the 32-column slice of CTA s8
              const uint32_t rb = mapa(bar(BAR_F2), peer);
              const uint32_t base = mapa(sbase + Y::OFF_F2RECV + (rank * GMX * 32 + (feat & 31)) * 4, peer);
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                const int i0 = 8 * nb + 2 * fq;
                if (i0 < G) { st_async_b32(base + i0 * 128, __float_as_uint(acc[nb][0]), rb); st_async_b32(base + i0 * 128 + 32, __float_as_uint(acc[nb][2]), rb); }
                if (i0 + 1 < G) { st_async_b32(base + (i0 + 1) * 128, __float_as_uint(acc[nb][1]), rb); st_async_b32(base + (i0 + 1) * 128 + 32, __float_as_uint(acc[nb][3]), rb); }
              }
              stage_release(4);
            } else stage_skip();
          }
          TRACE(t);   // 13: FFN2 issued
          {  // reduce the 8 partial slices of the own 32 columns, add bias, all-gather, LN3
            const float b2 = __ldg(P.b_f2[l] + rank * 32 + c_t);
            if (tid == 0) mbar_expect_tx(bar(BAR_F2), G * DM * 4);
            xwait(BAR_F2, ph_f2);
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              const int img = gi_t + 8 * nb;
              float a = b2;
              if (img < G) {
#pragma unroll
                for (int p = 0; p < CS; ++p) a += f2recv[(p * GMX + img) * 32 + c_t];
              }
              push_y(img, a);
            }
          }
          TRACE(t);   // 14: FFN2 reduced + pushed
          layer_norm(P.ln3w[l], P.ln3b[l]);
          TRACE(t);   // 15: LN3

        }
        // ---- vocabulary head: own 40 rows -> logits to the caller's tensor and to the CTA that selects for the image ---
        const bool want_conf = P.confs && (t % 4 == 0);
        const bool need_select = !P.forced || want_conf;
        {
          const int r0 = rank * VSL;
          float b0 = 0.f, b1 = 0.f;
          const int row_a = warp * 16 + fg, row_b = row_a + 8;
          const bool va = warp < 3 && row_a < VSL && r0 + row_a < P.vocab, vb = warp < 3 && row_b < VSL && r0 + row_b < P.vocab;
          if (va) b0 = __ldg(P.b_out + r0 + row_a);
          if (vb) b1 = __ldg(P.b_out + r0 + row_b);
          // rows [0,32) of the own 40 sit in one stage (warps 0-1), rows [32,40) in the next (warp 2: an 8-row block, the upper half
          // of its 16-row MMA tile reads the neighbouring k-block -- finite weights feeding accumulator rows that are never used)
          auto head_out = [&](float (&acc)[NB][4]) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int img = 8 * nb + 2 * fq + (e & 1); const bool hi8 = e >= 2;
                const bool valid = (hi8 ? vb : va) && img < G;
                if (valid) {
                  const int row = hi8 ? row_b : row_a;
                  const float lg = acc[nb][e] + (hi8 ? b1 : b0);
                  if (P.logits_out) P.logits_out[(int64_t)(img0 + img) * P.logits_img_stride + (int64_t)(t + P.logits_row_offset) * P.vocab + r0 + row] = lg;
                  // the CTA that selects for image img is img % 8; it keeps one [CS][VSL] receive block per owned image
                  if (need_select) st_async_b32(mapa(sbase + Y::OFF_LRECV + ((nb * CS + rank) * VSL + row) * 4, img & 7), __float_as_uint(lg), mapa(bar(BAR_LG), img & 7));
                }
              }
          };
          if (warp < 2) {
            const uint32_t st = stage_wait();
            float acc[NB][4];
            mma_mtile<32, NB>(st, warp * 16, sbase + Y::OFF_XH, acc);
            head_out(acc);
            stage_release(4);
          } else stage_skip();
          if (warp == 2) {
            const uint32_t st = stage_wait();
            float acc[NB][4];
            mma_mtile<8, NB>(st, 0, sbase + Y::OFF_XH, acc);
            head_out(acc);
            stage_release(8);
          } else stage_skip();
        }
        cbar();       // the next step's embedding overwrites the operand rows the head MMAs of warps 0-2 are reading (in teacher-forced
                      // mode nothing else orders the two: no select, no token exchange)
        TRACE(t);     // head done
        // ---- select: CTA `rank` owns images `rank` (and `rank + 8`) ---------------------------------------------------
        if (need_select && (int)rank < G) {
          const int V = P.vocab;
          const int n_own = ((int)rank + 8 < G && NB > 1) ? 2 : 1;
          if (tid == 0) mbar_expect_tx(bar(BAR_LG), n_own * V * 4);
          mbar_wait(bar(BAR_LG), ph_lg);
          ph_lg ^= 1;
          auto publish = [&](int img, int token, float conf) {     // one thread: token to the caller's buffer and to every peer
            if (!P.forced) {
              P.tokens[(int64_t)(img0 + img) * P.tokens_ld + t + 1] = token;
              const uint32_t off = sbase + Y::OFF_TOK + ((((t + 1) & 1) * GMX) + img) * 4;
#pragma unroll
              for (int p = 0; p < CS; ++p) st_async_b32(mapa(off, p), (uint32_t)token, mapa(bar(BAR_TOK), p));
            }
            if (want_conf) P.confs[(int64_t)(img0 + img) * P.confs_ld + t / 4] = conf;
          };
          if (P.top_k == 0 && P.top_p == 1.0f) {
            // greedy: one warp per owned image straight from the receive buffer ([src][VSL] is vocabulary order), no block barrier
            if (warp < n_own) {
              int token; float conf;
              warp_greedy_select(lrecv + warp * CS * VSL, V, token, conf);
              if (lane == 0) publish(rank + 8 * warp, token, conf);
            }
          } else {
            int Vp2 = 1; while (Vp2 < V) Vp2 <<= 1;
            float* lg = selbuf; float* srt = selbuf + CS * VSL;
            const bool sample = P.uniforms != nullptr;
            for (int nb = 0; nb < n_own; ++nb) {
              const int img = rank + 8 * nb;
              for (int i = tid; i < V; i += NCT) lg[i] = lrecv[nb * CS * VSL + i];
              cbar();
              const float u = sample ? P.uniforms[(int64_t)(img0 + img) * P.uniforms_ld + t] : 0.f;
              int token; float conf;
              select_from_logits(lg, srt, V, Vp2, P.top_k, P.top_p, sample, u, token, conf);
              if (tid == 0) publish(img, token, conf);
              if (nb + 1 < n_own) cbar();      // the scratch is reused by the second image
            }
          }
        }
        if (!P.forced) {
          if (tid == 0) mbar_expect_tx(bar(BAR_TOK), G * 4);
          xwait(BAR_TOK, ph_tok);
        }
        TRACE(t);     // tokens exchanged
      }
      if (kTrace && tid == 0 && blockIdx.x == 0) { P.trace[200] = cons_wait; P.trace[201] = exch_wait; }
    }
  }
  // no CTA may exit while a peer can still push into its shared memory
  __syncthreads();
  cluster_sync_all();
}

// ---- host side: cached weight tensor maps ---------------------------------------------------------------------
struct FusedCache {
  const void* key[8 * 6 + 1];
  FusedParams P;
  bool valid;
};

// ---- kernel variants --------------------------------------------------------------------------------------------------
// 0: NB = 1, 10-stage ring (160 KB), one CTA per SM   (<= 8 images per cluster; lowest latency of one batch)
// 1: NB = 2, 8-stage ring (128 KB), one CTA per SM    (<= 16 images per cluster)
// 2: NB = 1, 4-stage ring (64 KB), compact layout, two CTAs per SM (<= 8 images per cluster; two clusters interleave on the same SMs)
constexpr int N_VARIANTS = 3;
template <bool kTrace> struct Variants {
  static const void* fn(int v) {
    switch (v) {
      case 0: return (const void*)decode_fused_kernel<kTrace, 1, 10, 1>;
      case 1: return (const void*)decode_fused_kernel<kTrace, 2, 8, 1>;
      default: return (const void*)decode_fused_kernel<kTrace, 1, 4, 2>;
    }
  }
};
int variant_smem(int v) { return v == 0 ? Lay<1, 10, false>::SMEM_BYTES : (v == 1 ? Lay<2, 8, true>::SMEM_BYTES : Lay<1, 4, true>::SMEM_BYTES); }

// per-context (= per-device) launch state: the dynamic-smem opt-in and the occupancy query are device properties
struct ClusterCtxState { int max_clusters[N_VARIANTS]; };

}  // namespace

int decode_cluster_supported(const mdc_model* m, const mdc_decode_state* st, int t_end) {
  const mdc_dims& d = m->d;
  if (d.precision != MDC_BF16 || d.dec_loop_dtype != MDC_F16 || d.dim != DM || d.dec_heads != CS || d.dec_ffn != FFN) return 0;
  if (d.dec_layers < 1 || d.dec_layers > 8 || d.vocab > CS * VSL || d.vocab < 8) return 0;
  if (st->x_override || st->pos_override || st->per_op_kernels) return 0;
  if (d.n_patches > 256 || d.n_patches < 8) return 0;             // one TMA box (<= 256 rows) per image panel, 2 panels per stage
  if (t_end > 256 || st->pages_per_seq > 32 || d.page_tokens != 16 || st->n_pages < 1) return 0;
  return 1;
}

size_t decode_cluster_scratch_bytes(const mdc_model*, int) { return 0; }

void decode_cluster_model_destroy(mdc_model* m) {
  if (m && m->fused_cache) { delete (FusedCache*)m->fused_cache; m->fused_cache = nullptr; }
}

void decode_cluster_ctx_destroy(mdc_ctx* ctx) {
  if (ctx && ctx->decode_state) { delete (ClusterCtxState*)ctx->decode_state; ctx->decode_state = nullptr; }
}

int decode_cluster_launch(mdc_model* m, const mdc_decode_state* st, int t_begin, int t_end, void* /*logits_scratch*/, cudaStream_t s) {
  mdc_ctx* ctx = m->ctx; const mdc_dims& d = m->d;
  const void** gw = m->w + MDC_ENC_GLOBAL_SLOTS + d.enc_depth * MDC_ENC_BLOCK_SLOTS;
  const void** lw0 = gw + MDC_DEC_GLOBAL_SLOTS;
  if (!m->fused_cache) { FusedCache* c = new FusedCache(); memset(c, 0, sizeof(FusedCache)); m->fused_cache = c; }
  FusedCache* fc = (FusedCache*)m->fused_cache;
  if (!fc->valid) {
    FusedParams& P = fc->P;
    for (int l = 0; l < d.dec_layers; ++l) {
      const void** lw = lw0 + l * MDC_DEC_LAYER_SLOTS;
      MDC_TRY(mdc_make_tmap_2d(ctx, lw[MDC_SA_IN_W], 3 * DM, DM, DM, 64, 32, 3, &P.m_in[l]));
      MDC_TRY(mdc_make_tmap_2d(ctx, lw[MDC_SA_OUT_W], DM, DM, DM, 64, 32, 3, &P.m_so[l]));
      MDC_TRY(mdc_make_tmap_2d(ctx, lw[MDC_CA_IN_W], 3 * DM, DM, DM, 64, 32, 3, &P.m_ca[l]));
      MDC_TRY(mdc_make_tmap_2d(ctx, lw[MDC_CA_OUT_W], DM, DM, DM, 64, 32, 3, &P.m_co[l]));
      MDC_TRY(mdc_make_tmap_2d(ctx, lw[MDC_FF1_W], FFN, DM, DM, 64, 32, 3, &P.m_f1[l]));
      MDC_TRY(mdc_make_tmap_2d(ctx, lw[MDC_FF2_W], DM, FFN, FFN, 64, 32, 3, &P.m_f2[l]));
      P.b_in[l] = (const float*)lw[MDC_SA_IN_B]; P.b_so[l] = (const float*)lw[MDC_SA_OUT_B];
      P.ln1w[l] = (const float*)lw[MDC_LN1_W]; P.ln1b[l] = (const float*)lw[MDC_LN1_B];
      P.b_ca[l] = (const float*)lw[MDC_CA_IN_B]; P.b_co[l] = (const float*)lw[MDC_CA_OUT_B];
      P.ln2w[l] = (const float*)lw[MDC_LN2_W]; P.ln2b[l] = (const float*)lw[MDC_LN2_B];
      P.b_f1[l] = (const float*)lw[MDC_FF1_B]; P.b_f2[l] = (const float*)lw[MDC_FF2_B];
      P.ln3w[l] = (const float*)lw[MDC_LN3_W]; P.ln3b[l] = (const float*)lw[MDC_LN3_B];
    }
    MDC_TRY(mdc_make_tmap_2d(ctx, gw[MDC_OUT_W], d.vocab, DM, DM, 64, 32, 3, &P.m_head));
    MDC_TRY(mdc_make_tmap_2d(ctx, gw[MDC_OUT_W], d.vocab, DM, DM, 64, 8, 3, &P.m_head8));
    P.emb = (const float*)gw[MDC_EMB]; P.pos = (const float*)gw[MDC_DEC_POS]; P.b_out = (const float*)gw[MDC_OUT_B];
    P.layers = d.dec_layers; P.vocab = d.vocab; P.S = d.n_patches; P.pad_idx = d.pad_idx; P.PT = d.page_tokens;
    fc->valid = true;
  }
  FusedParams P = fc->P;
  P.B = st->B;
  P.tokens = st->tokens; P.tokens_ld = st->tokens_ld;
  P.kv_pool = (bf16*)st->kv_pool; P.page_table = st->page_table; P.pages_per_seq = st->pages_per_seq;
  P.logits_out = st->logits; P.logits_img_stride = (int64_t)st->logits_ld * d.vocab; P.logits_row_offset = st->logits_row_offset;
  P.confs = st->confs; P.confs_ld = st->confs_ld;
  P.uniforms = st->uniforms; P.uniforms_ld = st->uniforms_ld; P.top_k = st->top_k; P.top_p = st->top_p; P.forced = st->forced;
  P.t_begin = t_begin; P.t_end = t_end;
  P.trace = nullptr; P.trace_t = -1;
  int ipc = st->images_per_cluster, cps = st->ctas_per_sm;
#ifdef MDC_DEVTOOLS   // developer build only (tools/decode_trace.py): phase-trace buffer and variant overrides from the environment
  if (const char* tp = getenv("MDC_DECODE_TRACE_PTR")) { P.trace = (long long*)strtoull(tp, nullptr, 0); const char* tt = getenv("MDC_DECODE_TRACE_T"); P.trace_t = tt ? atoi(tt) : t_begin; }
  if (const char* e = getenv("MDC_DECODE_IPC")) ipc = atoi(e);
  if (const char* e = getenv("MDC_DECODE_CPS")) cps = atoi(e);
#endif
  // cross-K/V [layers*B*S rows][2*DM]: one (S rows x 32 channels) box per (image, head, k|v); paged pool: one page x head box
  MDC_TRY(mdc_make_tmap_2d(ctx, st->cross_kv, (int64_t)d.dec_layers * st->B * d.n_patches, 2 * DM, 2 * DM, HD, d.n_patches, 2, &P.m_ckv));
  MDC_TRY(mdc_make_tmap_2d(ctx, st->kv_pool, (int64_t)st->n_pages * d.dec_layers * 2 * d.page_tokens, DM, DM, HD, d.page_tokens, 2, &P.m_pool));
  // variant: more than 8 images per cluster -> the two-column-block instantiation; otherwise ctas_per_sm == 2 -> the compact one
  const int variant = ipc > 8 ? 1 : (cps == 2 ? 2 : 0);
  const int gmx = variant == 1 ? 16 : 8;
  if (!ctx->decode_state) { ClusterCtxState* cs = new ClusterCtxState(); memset(cs, 0, sizeof(*cs)); ctx->decode_state = cs; }
  ClusterCtxState* cs = (ClusterCtxState*)ctx->decode_state;
  const int smem = variant_smem(variant);
  if (!cs->max_clusters[variant]) {
    const void* fn = Variants<false>::fn(variant);
    MDC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
#ifdef MDC_DEVTOOLS
    MDC_CUDA(cudaFuncSetAttribute(Variants<true>::fn(variant), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
#endif
    cudaLaunchConfig_t q{}; q.gridDim = dim3(CS * 64); q.blockDim = dim3(NT); q.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CS; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    q.attrs = a; q.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, fn, &q);
    if (e != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = ctx->sm_count / CS / 2; if (n < 1) n = 1; }
    cs->max_clusters[variant] = n;
  }
  const int max_clusters = cs->max_clusters[variant];
  int G = (P.B + max_clusters - 1) / max_clusters;
  if (ipc > 0 && ipc > G) G = ipc;                                                              // fewer, fuller clusters (batch pipelining)
  if (G > gmx) G = gmx;
  if (G < 1) G = 1;
  P.G = G;
  P.n_groups = (P.B + G - 1) / G;
  const int n_clusters = P.n_groups < max_clusters ? P.n_groups : max_clusters;
  void* args[1] = {(void*)&P};
  const void* fn = Variants<false>::fn(variant);
#ifdef MDC_DEVTOOLS
  if (P.trace) fn = Variants<true>::fn(variant);
#endif
  MDC_CUDA(cudaLaunchKernel(fn, dim3(n_clusters * CS), dim3(NT), args, (size_t)smem, s));
  MDC_LAUNCH_CHECK(ctx);
  return 0;
}
