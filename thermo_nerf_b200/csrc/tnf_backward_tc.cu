// Tensor-core backward of ThermalNerfactoTField (thermo_nerf/thermal_nerf/thermal_field.py:108-201, the autograd of
// :160-179) for sm_100a: ONE kernel per step, nothing per-sample written to HBM.
//
// CTA = 16 warps, one CTA per SM.  Warps 0..14 are producers: each owns one ray at a time - compositing backward
// (suffix scans), field MLPs recomputed from the fp16 hash features the forward saved, the dX chain on mma.sync
// (bf16 operands, fp32 accumulate), run-merged hash-grid scatter (REDG.F32x2).  Warp 15 is the weight-gradient
// issuer: the eight dW = dY^T X contractions (K = all samples of the batch: the one genuinely dense product of the
// step) run on the 5th-generation tensor cores with their 14 336 fp32 accumulators resident in tensor memory for
// the whole kernel - tcgen05.mma.cta_group::1.kind::f16, M = 64, K = 16 samples per instruction, both operands read
// from shared memory where the producers leave their (X, dY) tiles.  Registers could not hold those accumulators
// beside the per-sample chain (56 KB per accumulating agent), and shared-memory fp32 atomics are CAS loops on
// sm_100; tensor memory holds them for free and the single issuing thread costs no registers in the producers.
//
// Operand tiles: a [16 samples][W features] bf16 matrix is stored as W/8 column groups of 256 B, each group two
// 8x8 core matrices (samples 0-7, 8-15; 16 B per sample row) - the no-swizzle MN-major canonical layout of the
// tcgen05 shared-memory descriptor (leading byte offset = 128 B between the K halves, stride byte offset = 256 B
// between column groups; verified on hardware by scripts/probe_umma.cu).  One stmatrix.x4 writes four core
// matrices straight from the mma C-fragment registers, bank-conflict free.  Any run of adjacent groups is an
// operand, so [X | 1] (bias gradient as one more column) or [XG | 1 | SH | appearance] are just adjacent tiles.
//
// Producer -> issuer hand-off: a ring of kNumBufs (7) buffers of kBufGroups (31) 256-byte column groups = 7.75 KB; a producer takes a ticket (shared-memory
// atomic), waits until the ticket that used the buffer before has retired, fills it, fences generic -> async
// proxy and arrives on the buffer's "full" mbarrier.  The issuer consumes tickets in order, commits each to the
// buffer's "done" mbarrier (tcgen05.commit), observes those completions in order and publishes a monotonic
// `completed` counter - producers wait on the counter, never on a parity that could be two ring laps old.  A
// producer never holds a buffer while it waits for another one, so the ring cannot deadlock.  Long-lived operands
// (the trunk activation XH, the geo features XG and the per-ray [1 | SH | appearance] rows) live in a private
// per-producer area that is rewritten once the ticket of the tile's last product has retired.
//
// Epilogue: all 16 warps read the accumulators back (tcgen05.ld.32x32b) and add them to the global gradient
// tensors (one atomic per weight per CTA).
#include "tnf_backward.cuh"
#include "tnf_host.h"

namespace tnf {

template <int NT, int KT>
__device__ __forceinline__ void mma_layer_bf16(float (&c)[NT][4], const uint32_t (&a)[KT][4],
                                               const uint2* __restrict__ w, const int ntw, const int lane) {
  static_assert((NT & 1) == 0, "backward weight images use the paired-fragment layout");
  const uint4* __restrict__ w4 = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int nt = 0; nt < NT; nt += 2)
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const uint4 b = w4[(kt * (ntw >> 1) + (nt >> 1)) * 32 + lane];
      mma_16816_bf16(c[nt], a[kt], make_uint2(b.x, b.y));
      mma_16816_bf16(c[nt + 1], a[kt], make_uint2(b.z, b.w));
    }
}
template <int NT>
__device__ __forceinline__ void zero_c(float (&c)[NT][4]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f;
}
// C fragments -> bf16 A fragments of the next product (same index map as act_pack)
template <int NT>
__device__ __forceinline__ void pack_bf16_a(const float (&c)[NT][4], uint32_t (&a)[NT / 2][4]) {
#pragma unroll
  for (int kt = 0; kt < NT / 2; ++kt) {
    a[kt][0] = pack_bf162(c[2 * kt][0], c[2 * kt][1]);
    a[kt][1] = pack_bf162(c[2 * kt][2], c[2 * kt][3]);
    a[kt][2] = pack_bf162(c[2 * kt + 1][0], c[2 * kt + 1][1]);
    a[kt][3] = pack_bf162(c[2 * kt + 1][2], c[2 * kt + 1][3]);
  }
}
template <int NT>
__device__ __forceinline__ uint32_t relu_mask(float (&c)[NT][4]) {  // applies ReLU in place, returns the >0 mask
  uint32_t mk = 0;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c[nt][e] > 0.f) mk |= 1u << (nt * 4 + e); else c[nt][e] = 0.f;
    }
  return mk;
}
template <int NT>
__device__ __forceinline__ void apply_mask(float (&c)[NT][4], const uint32_t mk) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (!((mk >> (nt * 4 + e)) & 1u)) c[nt][e] = 0.f;
}

// bf16 B fragments of the backward products: B[kk][nn] = V(kk, nn), [kt][nt][lane]
struct FieldBwdWTC {
  uint2 rgb1T[4][8][32];   // dA1 = dA2pre . W_rgb1
  uint2 th1T[4][8][32];    // dB1 = dB2pre . W_th1
  uint2 geoT[8][2][32];    // dG  = [dA1pre | dB1pre] . [W_rgb0[:,16:31] ; W_th0]   (column 0 = density slot = 0)
  uint2 base1T[1][8][32];  // dH  = dG . W_base1
  uint2 base0T[4][4][32];  // dF  = dHpre . W_base0
  float rgb2w[4][64];      // row 3: the temperature channel of the RGBT head (concat_nerf), else zero
  float th2w[64];
};
struct ViewRows {  // V(k,n) = w[k*ld + n]
  const float* w;
  int ld, kv, nv;
  __device__ float operator()(int k, int n) const { return (k < kv && n < nv) ? w[k * ld + n] : 0.f; }
};
struct ViewGeoT {
  const float* rgb0;
  const float* th0;
  bool detach;
  __device__ float operator()(int k, int n) const {
    if (n < 1 || n > 15) return 0.f;
    if (k < 64) return rgb0[k * 63 + 16 + n - 1];
    return detach ? 0.f : th0[(k - 64) * 15 + n - 1];
  }
};
template <typename V>
__device__ inline void stage_frag_bf16(uint2* dst, int KT, int NT, const V& v, int tid, int nthreads) {
  for (int i = tid; i < KT * NT * 32; i += nthreads) {
    const int lane = i & 31, nt = (i >> 5) % NT, kt = (i >> 5) / NT;
    const int g = lane >> 2, q = lane & 3;
    const int n = nt * 8 + g, k = kt * 16 + 2 * q;
    dst[frag_index(kt, nt, lane, NT)] =
        make_uint2(pack_bf162(v(k, n), v(k + 1, n)), pack_bf162(v(k + 8, n), v(k + 9, n)));
  }
}
__device__ inline void stage_field_bwd(FieldBwdWTC& W, const TnfModel& m, int tid, int nthreads) {
  const TnfField& f = m.field;
  stage_frag_bf16(&W.rgb1T[0][0][0], 4, 8, ViewRows{f.rgb1.weight, 64, 64, 64}, tid, nthreads);
  stage_frag_bf16(&W.th1T[0][0][0], 4, 8, ViewRows{f.th1.weight, 64, 64, 64}, tid, nthreads);
  stage_frag_bf16(&W.geoT[0][0][0], 8, 2, ViewGeoT{f.rgb0.weight, f.th0.weight, m.detach_thermal_geo != 0}, tid,
                  nthreads);
  stage_frag_bf16(&W.base1T[0][0][0], 1, 8, ViewRows{f.base1.weight, 64, 16, 64}, tid, nthreads);
  stage_frag_bf16(&W.base0T[0][0][0], 4, 4, ViewRows{f.base0.weight, 32, 64, 32}, tid, nthreads);
  const int nout = m.head_mode == TNF_HEAD_CONCAT ? 4 : 3;
  for (int i = tid; i < 256; i += nthreads) W.rgb2w[i / 64][i % 64] = i < 64 * nout ? f.rgb2.weight[i] : 0.f;
  for (int i = tid; i < 64; i += nthreads) W.th2w[i] = f.th2.weight[i];
}

// ------------------------------------------------------------------------------------
// tcgen05 / mbarrier plumbing
// ------------------------------------------------------------------------------------
constexpr int kProducers = 15;            // producer warps per CTA; warp kProducers issues the tcgen05.mma
constexpr int kBwdThreads = (kProducers + 1) * 32;
constexpr int kGrp = 256;                 // bytes of one 8-column group of a 16-sample bf16 tile
constexpr int kBufGroups = 31;            // largest event (EV_TRUNK): [XF 4 | 1 | dG 2 | dH 8 | dB1 8 | dA1 8]
constexpr int kNumBufs = 7;
constexpr int kTmemCols = 512;

// tensor-memory columns of the accumulators (all M = 64: row m on lane (m % 16) + 32 * (m / 16))
enum {
  C_TH1 = 0,      // [n][k | bias]      dB2^T [XB1 | 1]            72 columns
  C_RGB1 = 72,    // [n][k | bias]      dA2^T [XA1 | 1]            72
  C_RGB0 = 144,   // [n][geo16 | bias8 | sh16 | app32]  dA1^T [XG | 1 | SH | appearance]   72
  C_BASE0 = 216,  // [n][k | bias]      dH^T [XF | 1]              40
  C_TH0 = 256,    // [n][geo16 | bias8] dB1^T [XG | 1]             24
  C_BASE1 = 280,  // [k][n]             XH^T dG                    16
  C_TH2 = 296,    // [k][0]             XB2^T dT                    8
  C_RGB2 = 304,   // [k][c]             XA2^T dZ                    8
  C_SB = 312,     // row sums: [dT 8 | dZ 8 | dG 16] against an all-ones A tile   32
  C_END = 344
};
// One hand-off per head and one for the trunk: three events per 16-sample tile.
//   EV_THERMAL / EV_COLOUR  [X1 8 | 1 | dY2 8 | X2 8 | dOut 1]   layers 1 and 2 of the head
//   EV_TRUNK                [XF 4 | 1 | dG 2 | dH 8 | dB1 8 | dA1 8]  mlp_base + layer 0 of both heads
enum { EV_THERMAL = 0, EV_COLOUR, EV_TRUNK, EV_COUNT };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// non-blocking probe of a phase; the issuer observes every barrier in ticket order, one phase at a time, so the
// parity never aliases
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// The `completed` counter: written by the issuer after it has observed the commit barrier of a ticket (the
// tensor core has finished reading that buffer by then; stores are not speculated, so a plain volatile store
// behind the dependent branch suffices - st.release would add a MEMBAR.ALL.CTA to every retirement), read by
// producers before they overwrite the buffer with generic stores.
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;\n" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
constexpr long long kSpinLimit = 1ll << 32;  // ~2 s: a protocol error traps (the launch fails loudly), never hangs
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor of a run of column groups: no swizzle, K halves 128 B apart, groups 256 B apart
__device__ __forceinline__ uint64_t umma_desc(const uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(kGrp >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor: D f32, A / B bf16, both MN-major (sample-major tiles), M = 64, N = n
__device__ __forceinline__ uint32_t umma_idesc(const int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((64u >> 4) << 24);
}
// D[64 x n] (tensor-memory column `col`) (+)= A^T B over the 16 samples of the tiles at a / b
__device__ __forceinline__ void umma(const uint32_t tmem, const int col, const uint32_t a, const uint32_t b, const int n,
                                     uint32_t& inited, const int acc_id) {
  const uint32_t accumulate = (inited >> acc_id) & 1u;
  inited |= 1u << acc_id;
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + (uint32_t)col),
      "l"(umma_desc(a)), "l"(umma_desc(b)), "r"(umma_idesc(n)), "r"(accumulate)
      : "memory");
}

// ------------------------------------------------------------------------------------
// operand tiles
// ------------------------------------------------------------------------------------
// four core matrices (samples 0-7 / 8-15 of column groups grp, grp + 1) from fragment registers
__device__ __forceinline__ void stmatrix_groups(unsigned char* tile, const int grp, const uint32_t r0, const uint32_t r1,
                                                const uint32_t r2, const uint32_t r3, const int lane) {
  const int mi = lane >> 3;  // lane i addresses row (i & 7) of matrix (i >> 3)
  const uint32_t addr = smem_u32(tile + (grp + (mi >> 1)) * kGrp + (mi & 1) * 128 + (lane & 7) * 16);
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3)
               : "memory");
}
// mma C fragments (16 samples x NT*8 columns, fp32) -> bf16 tile
template <int NT>
__device__ __forceinline__ void stage_c(unsigned char* tile, const float (&c)[NT][4], const int lane) {
#pragma unroll
  for (int nt = 0; nt < NT; nt += 2)
    stmatrix_groups(tile, nt, pack_bf162(c[nt][0], c[nt][1]), pack_bf162(c[nt][2], c[nt][3]),
                    pack_bf162(c[nt + 1][0], c[nt + 1][1]), pack_bf162(c[nt + 1][2], c[nt + 1][3]), lane);
}
// mma A fragments (bf16) -> tile
template <int KT>
__device__ __forceinline__ void stage_a(unsigned char* tile, const uint32_t (&a)[KT][4], const int lane) {
#pragma unroll
  for (int kt = 0; kt < KT; ++kt) stmatrix_groups(tile, 2 * kt, a[kt][0], a[kt][1], a[kt][2], a[kt][3], lane);
}
__device__ __forceinline__ uint32_t half2_to_bf162(const uint32_t h) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
  return pack_bf162(f.x, f.y);
}
// mma A fragments holding fp16 (the forward activations) -> bf16 tile (one kind::f16 instruction takes one format)
template <int KT>
__device__ __forceinline__ void stage_a_f16(unsigned char* tile, const uint32_t (&a)[KT][4], const int lane) {
#pragma unroll
  for (int kt = 0; kt < KT; ++kt)
    stmatrix_groups(tile, 2 * kt, half2_to_bf162(a[kt][0]), half2_to_bf162(a[kt][1]), half2_to_bf162(a[kt][2]),
                    half2_to_bf162(a[kt][3]), lane);
}
// one column group whose rows are [1, 0, 0, 0, 0, 0, 0, 0]: the bias-gradient column of a B operand
__device__ __forceinline__ void stage_ones(unsigned char* grp, const int lane) {
  if (lane < 16) *reinterpret_cast<uint4*>(grp + (lane >> 3) * 128 + (lane & 7) * 16) = make_uint4(0x00003F80u, 0u, 0u, 0u);
}
// one column group from per-row values held by the q == 0 lane of rows g / g + 8
__device__ __forceinline__ void stage_rows(unsigned char* grp, const int g, const int q, const uint32_t lo0,
                                           const uint32_t hi0, const uint32_t lo1, const uint32_t hi1) {
  if (q == 0) {
    *reinterpret_cast<uint4*>(grp + g * 16) = make_uint4(lo0, hi0, 0u, 0u);
    *reinterpret_cast<uint4*>(grp + 128 + g * 16) = make_uint4(lo1, hi1, 0u, 0u);
  }
}

struct ProducerScratch {
  FieldBwdScratch s;
  float appv[32];  // appearance embedding of the ray's camera
};

struct alignas(1024) FieldBwdSmemU {
  unsigned char pool[kNumBufs][kBufGroups * kGrp];
  unsigned char xh[kProducers][8 * kGrp];    // trunk activation of the tile in flight (A operand of base1)
  unsigned char xgv[kProducers][9 * kGrp];   // [XG 2 groups | 1 | SH 2 | appearance 4] (B operand of th0 / rgb0)
  unsigned char ones[8 * kGrp];              // all-ones [16][64] A tile: row sums of the narrow dY matrices
  FieldWTC fw;
  FieldBwdWTC bw;
  ProducerScratch ws[kProducers];
  unsigned long long full[kNumBufs];   // producer -> issuer: buffer filled
  unsigned long long done[kNumBufs];   // tensor core -> issuer: the MMAs that read the buffer retired
  uint32_t hdr[kNumBufs];              // event type | producer warp << 8
  uint32_t ticket;                     // next ticket (producers, atomic)
  uint32_t completed;                  // tickets [0, completed) have retired (issuer, release / acquire)
  uint32_t tmem_base;
};
static_assert(sizeof(FieldBwdSmemU) <= 227 * 1024, "FieldBwdSmemU exceeds the shared memory of an sm_100 SM");

// Tickets are handed out in arrival order; ticket t uses buffer t % kNumBufs once ticket t - kNumBufs has retired.
// Producers wait on the monotonic `completed` counter (a parity wait could be two ring laps behind and alias).
// (No time-out here: the issuer traps when it sees no progress for kSpinLimit cycles, which ends the launch.)
__device__ __forceinline__ void wait_retired(FieldBwdSmemU& S, const uint32_t upto) {  // until completed >= upto
  while ((int32_t)(ld_volatile(&S.completed) - upto) < 0) __nanosleep(20);
}
__device__ __forceinline__ uint32_t acquire_ticket(FieldBwdSmemU& S, const int lane) {
  uint32_t t = 0;
  if (lane == 0) {
    t = atomicAdd(&S.ticket, 1u);
    if (t >= (uint32_t)kNumBufs) wait_retired(S, t - kNumBufs + 1);
  }
  return __shfl_sync(kFull, t, 0);
}
__device__ __forceinline__ void close_buf(FieldBwdSmemU& S, const uint32_t ticket, const int ev, const int warp,
                                          const int lane) {
  const int b = (int)(ticket % kNumBufs);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> tensor-core reads
  __syncwarp();
  if (lane == 0) {
    S.hdr[b] = (uint32_t)ev | ((uint32_t)warp << 8);
    mbar_arrive(&S.full[b]);
  }
}

// rays of this CTA: ray = (blockIdx.x + k * gridDim.x) * kProducers + warp
__device__ __forceinline__ long long cta_ray_count(const long long R) {
  long long n = 0;
  for (int w = 0; w < kProducers; ++w) {
    const long long slots = (R - w + kProducers - 1) / kProducers;  // slots s with s * kProducers + w < R
    if (slots > (long long)blockIdx.x) n += (slots - blockIdx.x + gridDim.x - 1) / gridDim.x;
  }
  return n;
}

// accumulator element (row m of the M = 64 tile, absolute column) -> gradient tensor
__device__ __forceinline__ void store_grad(const TnfFieldGrad& g, const int col, const int row, const float v,
                                           const int nout) {
  if (v == 0.f) return;
  if (col < C_RGB1) {
    const int c = col - C_TH1;
    if (c < 64) atomicAdd(g.th1.weight + row * 64 + c, v);
    else if (c == 64) atomicAdd(g.th1.bias + row, v);
  } else if (col < C_RGB0) {
    const int c = col - C_RGB1;
    if (c < 64) atomicAdd(g.rgb1.weight + row * 64 + c, v);
    else if (c == 64) atomicAdd(g.rgb1.bias + row, v);
  } else if (col < C_BASE0) {  // mlp_head.layers.0: input = [SH 16 | geo 15 | appearance 32]
    const int c = col - C_RGB0;
    if (c >= 1 && c <= 15) atomicAdd(g.rgb0.weight + row * 63 + 16 + c - 1, v);
    else if (c == 16) atomicAdd(g.rgb0.bias + row, v);
    else if (c >= 24 && c < 40) atomicAdd(g.rgb0.weight + row * 63 + (c - 24), v);
    else if (c >= 40) atomicAdd(g.rgb0.weight + row * 63 + 31 + (c - 40), v);
  } else if (col < C_TH0) {
    const int c = col - C_BASE0;
    if (c < 32) atomicAdd(g.base0.weight + row * 32 + c, v);
    else if (c == 32) atomicAdd(g.base0.bias + row, v);
  } else if (col < C_BASE1) {
    const int c = col - C_TH0;
    if (c >= 1 && c <= 15) atomicAdd(g.th0.weight + row * 15 + c - 1, v);
    else if (c == 16) atomicAdd(g.th0.bias + row, v);
  } else if (col < C_TH2) {
    atomicAdd(g.base1.weight + (col - C_BASE1) * 64 + row, v);
  } else if (col < C_RGB2) {
    if (col == C_TH2) atomicAdd(g.th2.weight + row, v);
  } else if (col < C_SB) {
    const int c = col - C_RGB2;
    if (c < nout) atomicAdd(g.rgb2.weight + c * 64 + row, v);
  } else if (row == 0) {
    const int c = col - C_SB;
    if (c == 0) atomicAdd(g.th2.bias, v);
    else if (c >= 8 && c < 8 + nout) atomicAdd(g.rgb2.bias + (c - 8), v);
    else if (c >= 16) atomicAdd(g.base1.bias + (c - 16), v);
  }
}

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
template <bool POSE>  // POSE: also produce dL/d ray origins / directions (camera-optimiser path; re-gathers the table)
__global__ void __launch_bounds__(kBwdThreads, 1)
    tnf_backward_field_kernel_tc(const __grid_constant__ TnfModel m, const __grid_constant__ TnfRays rays,
                                 const __grid_constant__ TnfSaved sv, const __grid_constant__ TnfOutputGrads go,
                                 const __grid_constant__ TnfModelGrad gr) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FieldBwdSmemU& S = *reinterpret_cast<FieldBwdSmemU*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  stage_field(S.fw, m.field, tid, kBwdThreads, m.head_mode == TNF_HEAD_CONCAT ? 4 : 3);
  stage_field_bwd(S.bw, m, tid, kBwdThreads);
  for (int i = tid; i < 8 * 16; i += kBwdThreads)
    *reinterpret_cast<uint4*>(S.ones + i * 16) = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (tid == 0) {
    for (int b = 0; b < kNumBufs; ++b) {
      mbar_init(&S.full[b], 1);
      mbar_init(&S.done[b], 1);
    }
    S.ticket = 0;
    S.completed = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == kProducers) {  // the issuer warp owns the tensor-memory allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const int S2 = m.num_samples[TNF_NUM_PROP];
  const long long R = rays.num_rays;
  const int tiles_per_ray = (S2 + 15) / 16;
  const long long total_events = cta_ray_count(R) * tiles_per_ray * EV_COUNT;

  if (warp == kProducers) {
    // ================================================================ weight-gradient issuer
    if (lane == 0) {
      uint32_t inited = 0;
      const uint32_t ones = smem_u32(S.ones);
      const uint32_t total = (uint32_t)total_events;
      uint32_t t = 0, c = 0;  // next ticket to issue, next ticket to retire
      long long last = clock64();
      while (c < total) {
        // both probes are issued before either result is consumed: their latencies overlap
        const bool can_retire = c < t, can_issue = t < total;
        const bool retired = can_retire && mbar_test(&S.done[c % kNumBufs], (c / kNumBufs) & 1);
        const bool filled = can_issue && mbar_test(&S.full[t % kNumBufs], (t / kNumBufs) & 1);
        if (retired) st_volatile(&S.completed, ++c);
        if (filled) {
          const int b = (int)(t % kNumBufs);
          tc_fence_after();
          const uint32_t hdr = *reinterpret_cast<volatile uint32_t*>(&S.hdr[b]);
          const int ev = (int)(hdr & 0xffu), w = (int)(hdr >> 8);
          const uint32_t buf = smem_u32(S.pool[b]);
          if (ev == EV_TRUNK) {  // [XF 4 | 1 | dG 2 | dH 8 | dB1 8 | dA1 8], private XH and [XG | 1 | SH | app]
            const uint32_t xgv = smem_u32(S.xgv[w]);
            umma(tmem, C_BASE0, buf + 7 * kGrp, buf, 40, inited, 0);
            umma(tmem, C_BASE1, smem_u32(S.xh[w]), buf + 5 * kGrp, 16, inited, 1);
            umma(tmem, C_SB + 16, ones, buf + 5 * kGrp, 16, inited, 2);
            umma(tmem, C_TH0, buf + 15 * kGrp, xgv, 24, inited, 3);
            umma(tmem, C_RGB0, buf + 23 * kGrp, xgv, 72, inited, 4);
          } else {  // [X1 8 | 1 | dY2 8 | X2 8 | dOut 1]
            const bool th = ev == EV_THERMAL;
            umma(tmem, th ? C_TH1 : C_RGB1, buf + 9 * kGrp, buf, 72, inited, th ? 5 : 8);
            umma(tmem, th ? C_TH2 : C_RGB2, buf + 17 * kGrp, buf + 25 * kGrp, 8, inited, th ? 6 : 9);
            umma(tmem, th ? C_SB : C_SB + 8, ones, buf + 25 * kGrp, 8, inited, th ? 7 : 10);
          }
          tc_commit(&S.done[b]);
          ++t;
        }
        if (retired || filled) last = clock64();
        else if (clock64() - last > kSpinLimit) __trap();
      }
    }
    __syncwarp();
  } else {
    // ================================================================ producers: one ray per warp
    const FieldWTC& W = S.fw;
    const FieldBwdWTC& B = S.bw;
    FieldBwdScratch& ws = S.ws[warp].s;
    float* appv = S.ws[warp].appv;
    unsigned char* const xh = S.xh[warp];
    unsigned char* const xgv = S.xgv[warp];
    const int g = lane >> 2, q = lane & 3;
    const TnfHashGrid& grid = m.field.grid;
    const uint32_t mask = (1u << grid.log2_size) - 1u;
    float2* __restrict__ gtab = reinterpret_cast<float2*>(gr.field.table);
    const uint32_t* __restrict__ F = static_cast<const uint32_t*>(sv.field_features);  // [Ns][16] half2
    constexpr bool pose = POSE;
    uint32_t trunk_ticket = 0;  // ticket + 1 of this warp's last EV_TRUNK: the last reader of its private area

    for (long long ray = (long long)blockIdx.x * kProducers + warp; ray < R;
         ray += (long long)gridDim.x * kProducers) {
      RayCtx rc;
      rc.ox = __ldg(rays.origins + ray * 3 + 0);
      rc.oy = __ldg(rays.origins + ray * 3 + 1);
      rc.oz = __ldg(rays.origins + ray * 3 + 2);
      rc.dx = __ldg(rays.directions + ray * 3 + 0);
      rc.dy = __ldg(rays.directions + ray * 3 + 1);
      rc.dz = __ldg(rays.directions + ray * 3 + 2);
      rc.s_near = spacing_fn(rays.nears ? __ldg(rays.nears + ray) : m.near_plane);
      rc.s_far = spacing_fn(rays.fars ? __ldg(rays.fars + ray) : m.far_plane);
      for (int i = lane; i <= S2; i += 32) ws.bins[i] = sv.sdist[TNF_NUM_PROP][ray * (S2 + 1) + i];
      float sh[16], app_lane;
      ray_bias_and_inputs(m, rays, ray, rc, lane, ws.rayb, sh, app_lane);
      appv[lane] = app_lane;
      __syncwarp();
      composite_backward(m, ws, rc, S2, lane, sv.field_samples + ray * S2 * 5,
                         go.weights[TNF_NUM_PROP] ? go.weights[TNF_NUM_PROP] + ray * S2 : nullptr,
                         go.rgb ? go.rgb[ray * 3 + 0] : 0.f, go.rgb ? go.rgb[ray * 3 + 1] : 0.f,
                         go.rgb ? go.rgb[ray * 3 + 2] : 0.f, go.thermal ? go.thermal[ray] : 0.f,
                         go.accumulation ? go.accumulation[ray] : 0.f);
      // the previous ray's last tile may still be read from the private area
      if (trunk_ticket) wait_retired(S, trunk_ticket);
      // per-ray rows [1 | SH | appearance] (identical for every sample): mlp_head.layers.0's direction /
      // appearance / bias gradients come out of the same product as its geo block
      if (lane < 16) {
        unsigned char* row = xgv + 2 * kGrp + (lane >> 3) * 128 + (lane & 7) * 16;
        *reinterpret_cast<uint4*>(row) = make_uint4(0x00003F80u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          *reinterpret_cast<uint4*>(row + (1 + j) * kGrp) =
              make_uint4(pack_bf162(sh[8 * j], sh[8 * j + 1]), pack_bf162(sh[8 * j + 2], sh[8 * j + 3]),
                         pack_bf162(sh[8 * j + 4], sh[8 * j + 5]), pack_bf162(sh[8 * j + 6], sh[8 * j + 7]));
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(row + (3 + j) * kGrp) =
              make_uint4(pack_bf162(appv[8 * j], appv[8 * j + 1]), pack_bf162(appv[8 * j + 2], appv[8 * j + 3]),
                         pack_bf162(appv[8 * j + 4], appv[8 * j + 5]), pack_bf162(appv[8 * j + 6], appv[8 * j + 7]));
      }
      float pg[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // pose path: sum_s dL/dx_s and sum_s t_s dL/dx_s of this ray
      ws.racc[lane] = 0.f;  // column sums of dA1pre over the ray (appearance-embedding gradient)
      ws.racc[lane + 32] = 0.f;

      for (int base = 0; base < S2; base += 16) {
        const int r0 = base + g, r1 = base + g + 8;
        const bool v0 = r0 < S2, v1 = r1 < S2;
        const int i0 = min(r0, S2 - 1), i1 = min(r1, S2 - 1);
        const long long row0 = ray * S2 + i0, row1 = ray * S2 + i1;
        float p[2][3], sel[2], mids[2];
        float dp[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};  // dL/d(normalised position) of rows r0 / r1 (pose path)
        {
          float delta;
          sample_geometry(rc, ws.bins[i0], ws.bins[i0 + 1], mids[0], delta);
          sel[0] = normalise_position(m, ray_x(rc, mids[0]), ray_y(rc, mids[0]), ray_z(rc, mids[0]), p[0][0], p[0][1], p[0][2]);
          sample_geometry(rc, ws.bins[i1], ws.bins[i1 + 1], mids[1], delta);
          sel[1] = normalise_position(m, ray_x(rc, mids[1]), ray_y(rc, mids[1]), ray_z(rc, mids[1]), p[1][0], p[1][1], p[1][2]);
        }
        // rows past the end of the ray carry zero upstream gradients: every dY row of theirs is zero, so the
        // (finite) duplicate X rows contribute nothing to the weight gradients
        const float dsig[2] = {v0 ? ws.dsig[i0] : 0.f, v1 ? ws.dsig[i1] : 0.f};
        // ws.dtau: thermal head -> dL/d thermal_i; RGBT head (concat_nerf) -> dL/d(pre-sigmoid channel 3), which
        // joins the colour head's output gradient while the thermal head sees none
        const bool concat = m.head_mode == TNF_HEAD_CONCAT;
        const float dt0 = v0 ? ws.dtau[i0] : 0.f, dt1 = v1 ? ws.dtau[i1] : 0.f;
        const float dtau[2] = {concat ? 0.f : dt0, concat ? 0.f : dt1};
        const float dz[2][4] = {{v0 ? ws.dzr[i0] : 0.f, v0 ? ws.dzg[i0] : 0.f, v0 ? ws.dzb[i0] : 0.f, concat ? dt0 : 0.f},
                                {v1 ? ws.dzr[i1] : 0.f, v1 ? ws.dzg[i1] : 0.f, v1 ? ws.dzb[i1] : 0.f, concat ? dt1 : 0.f}};
        // ---- saved hash features -> A fragments (fp16)
        uint32_t a0[2][4];
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
          for (int hl = 0; hl < 2; ++hl) {
            const int l = kt * 8 + hl * 4 + q;
            a0[kt][2 * hl] = __ldg(F + row0 * 16 + l);
            a0[kt][2 * hl + 1] = __ldg(F + row1 * 16 + l);
          }
        // the previous tile's weight-gradient products have read XH / XG
        if (trunk_ticket) wait_retired(S, trunk_ticket);
        // ---- trunk forward: H, G
        uint32_t hid[4][4];
        uint32_t mkH;
        {
          float c[8][4];
          init_bias(c, W.base0b, q);
          mma_layer<8, 2>(c, a0, &W.base0[0][0][0], 0, 8, lane);
          mkH = relu_mask(c);
          stage_c<8>(xh, c, lane);
          act_pack<8, ACT_NONE>(c, hid);
        }
        float h0r0, h0r1;
        uint32_t ga[1][4];
        {
          float c[2][4];
          init_bias(c, W.base1b, q);
          mma_layer<2, 4>(c, hid, &W.base1[0][0][0], 0, 2, lane);
          h0r0 = c[0][0];
          h0r1 = c[0][2];
          if (q == 0) { c[0][0] = 0.f; c[0][2] = 0.f; }
          stage_c<2>(xgv, c, lane);
          act_pack<2, ACT_NONE>(c, ga);
        }
        // ---- thermal head forward + backward down to dB1pre
        uint32_t dB1A[4][4];
        {
          float c[8][4];
          uint32_t xb1[4][4], dy[4][4];
          init_bias(c, W.th0b, q);
          mma_layer<8, 1>(c, ga, &W.geo0[0][0][0], 8, 16, lane);
          const uint32_t mkB1 = relu_mask(c);
          act_pack<8, ACT_NONE>(c, xb1);
          init_bias(c, W.th1b, q);
          mma_layer<8, 4>(c, xb1, &W.th1[0][0][0], 0, 8, lane);
          // x2 <- XB2 = sigmoid(.);  dy <- dB2pre = dtau w_th2 XB2 (1 - XB2)  (bf16 A fragments of the next product).
          // Everything the hand-off stores is packed BEFORE the ticket is taken: between acquire and close only the
          // stores remain (the issuer consumes tickets in order - a long fill would hold up every later ticket)
          uint32_t x2[4][4];
#pragma unroll
          for (int kt = 0; kt < 4; ++kt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int nt = 2 * kt + h;
              float x[4], d[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                x[e] = sigmoid_fast(c[nt][e]);
                d[e] = dtau[e >> 1] * B.th2w[nt * 8 + 2 * q + (e & 1)] * x[e] * (1.f - x[e]);
              }
              x2[kt][2 * h] = pack_bf162(x[0], x[1]);
              x2[kt][2 * h + 1] = pack_bf162(x[2], x[3]);
              dy[kt][2 * h] = pack_bf162(d[0], d[1]);
              dy[kt][2 * h + 1] = pack_bf162(d[2], d[3]);
            }
#pragma unroll
          for (int kt = 0; kt < 4; ++kt)
#pragma unroll
            for (int j = 0; j < 4; ++j) xb1[kt][j] = half2_to_bf162(xb1[kt][j]);
          {  // mlp_thermal.layers.1: dW = dB2^T [XB1 | 1];  field_head_thermal: dW^T = XB2^T dT
            const uint32_t d0 = pack_bf162(dtau[0], 0.f), d1 = pack_bf162(dtau[1], 0.f);
            const uint32_t tk = acquire_ticket(S, lane);
            unsigned char* buf = S.pool[tk % kNumBufs];
            stage_a<4>(buf, xb1, lane);
            stage_ones(buf + 8 * kGrp, lane);
            stage_a<4>(buf + 9 * kGrp, dy, lane);
            stage_a<4>(buf + 17 * kGrp, x2, lane);
            stage_rows(buf + 25 * kGrp, g, q, d0, 0u, d1, 0u);
            close_buf(S, tk, EV_THERMAL, warp, lane);
          }
          zero_c(c);
          mma_layer_bf16<8, 4>(c, dy, &B.th1T[0][0][0], 8, lane);
          apply_mask(c, mkB1);
          pack_bf16_a(c, dB1A);
        }
        // ---- colour head forward + backward down to dA1pre
        uint32_t dA1A[4][4];
        {
          float c[8][4];
          uint32_t xa1[4][4], dy[4][4];
          init_bias(c, ws.rayb, q);
          mma_layer<8, 1>(c, ga, &W.geo0[0][0][0], 0, 16, lane);
          const uint32_t mkA1 = relu_mask(c);
          act_pack<8, ACT_NONE>(c, xa1);
          init_bias(c, W.rgb1b, q);
          mma_layer<8, 4>(c, xa1, &W.rgb1[0][0][0], 0, 8, lane);
          // x2 <- XA2 = relu(.);  dy <- dA2pre = (XA2 > 0) dz . W_rgb2   (packed before the ticket, as above)
          uint32_t x2[4][4];
#pragma unroll
          for (int kt = 0; kt < 4; ++kt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int nt = 2 * kt + h;
              float x[4], d[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = nt * 8 + 2 * q + (e & 1), r = e >> 1;
                const float v = dz[r][0] * B.rgb2w[0][col] + dz[r][1] * B.rgb2w[1][col] + dz[r][2] * B.rgb2w[2][col] +
                                dz[r][3] * B.rgb2w[3][col];
                d[e] = c[nt][e] > 0.f ? v : 0.f;
                x[e] = fmaxf(c[nt][e], 0.f);
              }
              x2[kt][2 * h] = pack_bf162(x[0], x[1]);
              x2[kt][2 * h + 1] = pack_bf162(x[2], x[3]);
              dy[kt][2 * h] = pack_bf162(d[0], d[1]);
              dy[kt][2 * h + 1] = pack_bf162(d[2], d[3]);
            }
#pragma unroll
          for (int kt = 0; kt < 4; ++kt)
#pragma unroll
            for (int j = 0; j < 4; ++j) xa1[kt][j] = half2_to_bf162(xa1[kt][j]);
          {  // mlp_head.layers.1: dW = dA2^T [XA1 | 1];  mlp_head.layers.2: dW^T = XA2^T dZ
            const uint32_t z00 = pack_bf162(dz[0][0], dz[0][1]), z01 = pack_bf162(dz[0][2], dz[0][3]);
            const uint32_t z10 = pack_bf162(dz[1][0], dz[1][1]), z11 = pack_bf162(dz[1][2], dz[1][3]);
            const uint32_t tk = acquire_ticket(S, lane);
            unsigned char* buf = S.pool[tk % kNumBufs];
            stage_a<4>(buf, xa1, lane);
            stage_ones(buf + 8 * kGrp, lane);
            stage_a<4>(buf + 9 * kGrp, dy, lane);
            stage_a<4>(buf + 17 * kGrp, x2, lane);
            stage_rows(buf + 25 * kGrp, g, q, z00, z01, z10, z11);
            close_buf(S, tk, EV_COLOUR, warp, lane);
          }
          zero_c(c);
          mma_layer_bf16<8, 4>(c, dy, &B.rgb1T[0][0][0], 8, lane);
          apply_mask(c, mkA1);
          pack_bf16_a(c, dA1A);
          // column sums over the 16 rows of the tile -> per-ray sum of dA1pre
#pragma unroll
          for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              float sum = c[nt][b] + c[nt][b + 2];
              sum += __shfl_xor_sync(kFull, sum, 4);
              sum += __shfl_xor_sync(kFull, sum, 8);
              sum += __shfl_xor_sync(kFull, sum, 16);
              if (g == 0) ws.racc[nt * 8 + 2 * q + b] += sum;  // one lane per column: no conflicts, no atomics
            }
        }
        // ---- trunk backward: dG -> dH -> dF -> hash table
        uint32_t dGA[1][4];
        {
          float c[2][4];
          zero_c(c);
          mma_layer_bf16<2, 4>(c, dA1A, &B.geoT[0][0][0], 2, lane);
          mma_layer_bf16<2, 4>(c, dB1A, &B.geoT[4][0][0], 2, lane);
          if (q == 0) {  // density slot: trunc_exp backward times selector
            c[0][0] = dsig[0] * expf(fminf(fmaxf(h0r0, -15.f), 15.f)) * sel[0];
            c[0][2] = dsig[1] * expf(fminf(fmaxf(h0r1, -15.f), 15.f)) * sel[1];
          }
          pack_bf16_a(c, dGA);
        }
        uint32_t dHA[4][4];
        {
          float c[8][4];
          zero_c(c);
          mma_layer_bf16<8, 1>(c, dGA, &B.base1T[0][0][0], 8, lane);
          apply_mask(c, mkH);
          pack_bf16_a(c, dHA);
          {  // mlp_base: dW0 = dH^T [XF | 1], dW1^T = XH^T dG;  layer 0 of the heads: dB1^T [XG | 1],
             // dA1^T [XG | 1 | SH | appearance]
            // the tile's saved hash features (16 rows x 4 chunks of 8 halves) are fetched BEFORE the ticket is
            // taken: the issuer consumes tickets in order, so nothing slow may sit between acquire and close
            uint4 xf[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int idx = lane + 32 * i, row = idx >> 2, ch = idx & 3;
              const long long grow = ray * S2 + min(base + row, S2 - 1);
              xf[i] = __ldg(reinterpret_cast<const uint4*>(F + grow * 16) + ch);
              xf[i] = make_uint4(half2_to_bf162(xf[i].x), half2_to_bf162(xf[i].y), half2_to_bf162(xf[i].z),
                                 half2_to_bf162(xf[i].w));
            }
            const uint32_t tk = acquire_ticket(S, lane);
            unsigned char* buf = S.pool[tk % kNumBufs];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int idx = lane + 32 * i, row = idx >> 2, ch = idx & 3;
              *reinterpret_cast<uint4*>(buf + ch * kGrp + (row >> 3) * 128 + (row & 7) * 16) = xf[i];
            }
            stage_ones(buf + 4 * kGrp, lane);
            stage_a<1>(buf + 5 * kGrp, dGA, lane);
            stage_a<4>(buf + 7 * kGrp, dHA, lane);
            stage_a<4>(buf + 15 * kGrp, dB1A, lane);
            stage_a<4>(buf + 23 * kGrp, dA1A, lane);
            close_buf(S, tk, EV_TRUNK, warp, lane);
            trunk_ticket = tk + 1;
          }
        }
        {
          float c[4][4];
          zero_c(c);
          mma_layer_bf16<4, 4>(c, dHA, &B.base0T[0][0][0], 4, lane);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int l = nt * 4 + q;
            float2* lt = gtab + ((size_t)l << grid.log2_size);
            const float sc = W.scal[l];
            scatter_level_runs<4>(lt, p[0][0], p[0][1], p[0][2], sc, mask, v0 ? c[nt][0] : 0.f, v0 ? c[nt][1] : 0.f, lane);
            scatter_level_runs<4>(lt, p[1][0], p[1][1], p[1][2], sc, mask, v1 ? c[nt][2] : 0.f, v1 ? c[nt][3] : 0.f, lane);
            if (pose) {
              const float2* rt = reinterpret_cast<const float2*>(grid.table) + ((size_t)l << grid.log2_size);
              if (v0) hash_level_pos_grad(rt, p[0][0], p[0][1], p[0][2], sc, mask, c[nt][0], c[nt][1], dp[0][0], dp[0][1], dp[0][2]);
              if (v1) hash_level_pos_grad(rt, p[1][0], p[1][1], p[1][2], sc, mask, c[nt][2], c[nt][3], dp[1][0], dp[1][1], dp[1][2]);
            }
          }
        }
        if (pose) {
          // the four q-lanes of a row hold the contributions of their four levels each
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              dp[h][j] += __shfl_xor_sync(kFull, dp[h][j], 1);
              dp[h][j] += __shfl_xor_sync(kFull, dp[h][j], 2);
            }
            if (q == 0 && (h ? v1 : v0)) {
              position_grad_to_world(m, ray_x(rc, mids[h]), ray_y(rc, mids[h]), ray_z(rc, mids[h]), sel[h], dp[h][0],
                                     dp[h][1], dp[h][2]);
              pg[0] += dp[h][0]; pg[1] += dp[h][1]; pg[2] += dp[h][2];
              pg[3] = fmaf(mids[h], dp[h][0], pg[3]); pg[4] = fmaf(mids[h], dp[h][1], pg[4]);
              pg[5] = fmaf(mids[h], dp[h][2], pg[5]);
            }
          }
        }
      }
      if (pose) flush_ray_grad(gr, ray, pg[0], pg[1], pg[2], pg[3], pg[4], pg[5], lane);
      // ---- per-ray epilogue: appearance-embedding gradient = W_app^T sum_s dA1pre
      __syncwarp();
      if (m.appearance_mode == TNF_APPEARANCE_LOOKUP && gr.field.appearance) {
        const float* Wr = m.field.rgb0.weight;
        float acc = 0.f;
        for (int n = 0; n < 64; ++n) acc = fmaf(ws.racc[n], Wr[n * 63 + 31 + lane], acc);
        atomicAdd(gr.field.appearance + rays.camera_indices[ray] * 32 + lane, acc);
      }
      __syncwarp();
    }
  }

  // ================================================================ accumulators -> global gradients
  tc_fence_before();
  __syncthreads();
  if (total_events > 0) {  // the issuer left its loop after the last ticket retired
    tc_fence_after();
    const int qd = warp & 3;  // a warp reaches the 32 tensor-memory lanes of its quadrant
    for (int it = warp >> 2; it < C_END / 8; it += 4) {
      // CTAs finish together and add to the same 14 336 addresses: each starts at a different column chunk so
      // that the L2 atomic units do not see 148 adds to one address at the same moment
      const int chunk = (it + (int)blockIdx.x) % (C_END / 8);
      uint32_t r[8];
      const uint32_t taddr = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(chunk * 8);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (lane < 16) {  // M = 64: rows 16 qd .. 16 qd + 15 sit on the first 16 lanes of the quadrant
#pragma unroll
        for (int j = 0; j < 8; ++j)
          store_grad(gr.field, chunk * 8 + j, qd * 16 + lane, __uint_as_float(r[j]), m.head_mode == TNF_HEAD_CONCAT ? 4 : 3);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kProducers)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

int launch_backward_field_tc(const TnfModel& m, const TnfRays& rays, const TnfSaved& sv, const TnfOutputGrads& go,
                             const TnfModelGrad& gr, cudaStream_t stream) {
  const size_t smem = sizeof(FieldBwdSmemU);
  const long long want = (rays.num_rays + kProducers - 1) / kProducers;
  const long long cap = num_sms();
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaError_t e;
  if (gr.ray_origins) {
    e = cudaFuncSetAttribute(tnf_backward_field_kernel_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) tnf_backward_field_kernel_tc<true><<<grid, kBwdThreads, smem, stream>>>(m, rays, sv, go, gr);
  } else {
    e = cudaFuncSetAttribute(tnf_backward_field_kernel_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) tnf_backward_field_kernel_tc<false><<<grid, kBwdThreads, smem, stream>>>(m, rays, sv, go, gr);
  }
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "backward_field_tc: cudaFuncSetAttribute(smem=%zu): %s", smem,
                                    cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TNF_ERR_CUDA, "backward_field_tc launch: %s", cudaGetErrorString(e));
  return TNF_OK;
}

}  // namespace tnf
