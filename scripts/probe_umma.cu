// Hardware probe for the tcgen05 primitives the fused weight-gradient path relies on (sm_100a).
//
//   D[M, N] (fp32, TMEM) (+)= A^T-style operands from shared memory, both MN-major, no swizzle:
//   a staged matrix tile is [16 samples (K)][W features (MN)] bf16 stored as W/8 column groups of 256 B, each
//   group = two 8x8 core matrices (samples 0-7, then 8-15), 16 bytes per sample row.
//
// Checks, against a host reference on small exact integers:
//   * which of the two stride fields of the shared-memory descriptor is the MN-group stride
//   * M = 128 (lane = row) and M = 64 (rows 16i..16i+15 -> lanes 32i..32i+15) accumulator layouts
//   * N in {8, 16, 24, 40, 72, 136}
//   * accumulate (enable_input_d) over several issues, tcgen05.commit -> mbarrier
//   * M = 64 accumulator placed at lane offset 16
//   * bf16 A with fp16 B in one kind::f16 instruction
//   * operands written with stmatrix by other warps (generic proxy -> async proxy fence)
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/probe_umma scripts/probe_umma.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

struct Case {
  int M, N;          // instruction shape
  int a_groups;      // column groups of A in smem (M / 8)
  int b_groups;      // N / 8
  int lbo, sbo;      // descriptor stride fields, bytes
  int issues;        // number of accumulating issues (each with K = 16)
  int lane_off;      // TMEM lane offset of D (0 or 16)
  int b_fp16;        // B operand holds fp16
  int use_stmatrix;  // operands written with stmatrix by warps 1..3
};

__host__ __device__ inline int a_val(int issue, int k, int m) { return ((k * 3 + m * 5 + issue * 7) % 7) - 3; }
__host__ __device__ inline int b_val(int issue, int k, int n) { return ((k * 2 + n * 3 + issue) % 5) - 2; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int b_fp16) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D format f32
  d |= 1u << 7;                    // A format bf16
  d |= (b_fp16 ? 0u : 1u) << 10;   // B format bf16 / f16
  d |= 1u << 15;                   // A MN-major
  d |= 1u << 16;                   // B MN-major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const uint32_t a = smem_u32(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

// address of element (k, col) of a blocked-8 staged tile that starts at `base`
__device__ __forceinline__ unsigned char* tile_addr(unsigned char* base, int k, int col) {
  return base + (col >> 3) * 256 + (k >> 3) * 128 + (k & 7) * 16 + (col & 7) * 2;
}

__global__ void __launch_bounds__(128) probe_kernel(const Case c, float* out /* [128 lanes][512 cols] */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* A = smem;                                  // issues x a_groups x 256 B
  unsigned char* B = smem + c.issues * c.a_groups * 256;    // issues x b_groups x 256 B
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // ---- operands
  if (!c.use_stmatrix) {
    for (int is = 0; is < c.issues; ++is) {
      for (int i = tid; i < 16 * c.a_groups * 8; i += 128) {
        const int k = i / (c.a_groups * 8), m = i % (c.a_groups * 8);
        *reinterpret_cast<__nv_bfloat16*>(tile_addr(A + is * c.a_groups * 256, k, m)) =
            __float2bfloat16((float)a_val(is, k, m));
      }
      for (int i = tid; i < 16 * c.b_groups * 8; i += 128) {
        const int k = i / (c.b_groups * 8), n = i % (c.b_groups * 8);
        unsigned char* p = tile_addr(B + is * c.b_groups * 256, k, n);
        if (c.b_fp16) *reinterpret_cast<__half*>(p) = __float2half((float)b_val(is, k, n));
        else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16((float)b_val(is, k, n));
      }
    }
  } else if (warp >= 1) {
    // warps 1..3 hold the 16 x W matrices as mma C fragments (row g / g+8, columns nt*8 + 2q, +1) and store
    // them with stmatrix.x4: matrices (rows 0-7, nt), (rows 8-15, nt), (rows 0-7, nt+1), (rows 8-15, nt+1)
    const int g = lane >> 2, q = lane & 3;
    for (int is = 0; is < c.issues; ++is) {
      for (int which = 0; which < 2; ++which) {
        const int groups = which ? c.b_groups : c.a_groups;
        unsigned char* base = which ? B + is * c.b_groups * 256 : A + is * c.a_groups * 256;
        for (int nt = (warp - 1) * 2; nt < groups; nt += 6) {
          uint32_t r[4];
          for (int j = 0; j < 4; ++j) {
            const int ntj = nt + (j >> 1), row = g + (j & 1) * 8, col = ntj * 8 + 2 * q;
            float v0 = 0.f, v1 = 0.f;
            if (ntj < groups) {
              v0 = which ? (float)b_val(is, row, col) : (float)a_val(is, row, col);
              v1 = which ? (float)b_val(is, row, col + 1) : (float)a_val(is, row, col + 1);
            }
            const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            r[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          // lane i addresses row (i & 7) of matrix (i >> 3)
          const int mi = lane >> 3, mr = lane & 7;
          const int ntm = nt + (mi >> 1);
          unsigned char* addr = base + (ntm < groups ? ntm : nt) * 256 + (mi & 1) * 128 + mr * 16;
          if (nt + 1 < groups) {
            asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};\n" ::"r"(smem_u32(addr)),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                         : "memory");
          } else {
            asm volatile("stmatrix.sync.aligned.m8n8.x2.shared.b16 [%0], {%1,%2};\n" ::"r"(smem_u32(addr)),
                         "r"(r[0]), "r"(r[1])
                         : "memory");
          }
        }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // ---- zero the whole allocation so that untouched lanes / columns read as a sentinel-free 0
  {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < 512; col += 8) {
      const uint32_t z = 0x7fc00000u;  // NaN marks "never written"
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(taddr + col),
                   "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // ---- issue from one lane of warp 3 (not the allocating warp)
  if (warp == 3 && lane == 0) {
    const uint32_t idesc = make_idesc(c.M, c.N, c.b_fp16);
    const uint32_t d = tmem + ((uint32_t)c.lane_off << 16) + 32;  // column offset 32: not at the allocation base
    for (int is = 0; is < c.issues; ++is) {
      const uint64_t da = make_desc(smem_u32(A + is * c.a_groups * 256), c.lbo, c.sbo);
      const uint64_t db = make_desc(smem_u32(B + is * c.b_groups * 256), c.lbo, c.sbo);
      const uint32_t acc = is > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // ---- dump every lane / column
  {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int col = 0; col < 512; col += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr + col));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int j = 0; j < 8; ++j) out[(size_t)tid * 512 + col + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static bool run_case(const Case& c, const char* name) {
  float* dout;
  CK(cudaMalloc(&dout, 128 * 512 * sizeof(float)));
  CK(cudaMemset(dout, 0, 128 * 512 * sizeof(float)));
  const size_t smem = (size_t)c.issues * (c.a_groups + c.b_groups) * 256 + 4096;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(c, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s LAUNCH FAILED: %s\n", name, cudaGetErrorString(e));
    exit(3);  // a sticky error: nothing after it would run
  }
  std::vector<float> out(128 * 512);
  CK(cudaMemcpy(out.data(), dout, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaFree(dout));
  // expected
  std::vector<float> ref((size_t)c.M * c.N, 0.f);
  for (int is = 0; is < c.issues; ++is)
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        float s = 0.f;
        for (int k = 0; k < 16; ++k) s += (float)(a_val(is, k, m) * b_val(is, k, n));
        ref[(size_t)m * c.N + n] += s;
      }
  // expected lane of row m
  auto lane_of = [&](int m) { return c.M == 128 ? m : (m % 16) + 32 * (m / 16) + c.lane_off; };
  int bad = 0, touched_outside = 0;
  for (int m = 0; m < c.M; ++m)
    for (int n = 0; n < c.N; ++n)
      if (out[(size_t)lane_of(m) * 512 + 32 + n] != ref[(size_t)m * c.N + n]) ++bad;
  std::vector<char> expect_lane(128, 0);
  for (int m = 0; m < c.M; ++m) expect_lane[lane_of(m)] = 1;
  for (int l = 0; l < 128; ++l)
    for (int col = 0; col < 512; ++col) {
      const bool inside = expect_lane[l] && col >= 32 && col < 32 + c.N;
      if (!inside && out[(size_t)l * 512 + col] == out[(size_t)l * 512 + col]) ++touched_outside;  // not NaN
    }
  printf("%-28s M=%3d N=%3d lbo=%3d sbo=%3d issues=%d laneoff=%2d bfp16=%d stm=%d : mismatches=%d written_outside=%d %s\n",
         name, c.M, c.N, c.lbo, c.sbo, c.issues, c.lane_off, c.b_fp16, c.use_stmatrix, bad, touched_outside,
         bad == 0 ? "OK" : "FAIL");
  if (bad && c.M <= 128) {
    // help decoding: where does row 0 / row 17 / col 0 land?
    for (int m : {0, 1, 17, 63}) {
      if (m >= c.M) continue;
      int hits = 0;
      for (int l = 0; l < 128 && hits < 3; ++l) {
        int match = 0;
        for (int n = 0; n < c.N; ++n) match += out[(size_t)l * 512 + 32 + n] == ref[(size_t)m * c.N + n];
        if (match == c.N) { printf("    row %d found on lane %d\n", m, l); ++hits; }
      }
    }
    printf("    lane0 cols 32..39: ");
    for (int j = 0; j < 8; ++j) printf("%g ", out[32 + j]);
    printf("| ref row0: ");
    for (int j = 0; j < 8 && j < c.N; ++j) printf("%g ", ref[j]);
    printf("\n");
  }
  return bad == 0;
}

int main() {
  int ok = 1;
  // which field is the MN-group stride?  (group stride 256 B, k-group stride 128 B)
  const bool v1 = run_case({128, 64, 16, 8, 128, 256, 1, 0, 0, 0}, "M128 lbo=k sbo=mn");
  if (!v1) {
    const bool v2 = run_case({128, 64, 16, 8, 256, 128, 1, 0, 0, 0}, "M128 lbo=mn sbo=k");
    printf("=> descriptor convention: %s\n", v2 ? "LBO = MN-group stride, SBO = K-group stride" : "NEITHER");
    if (!v2) return 1;
  } else {
    printf("=> descriptor convention: LBO = K-group stride, SBO = MN-group stride\n");
  }
  const int L = v1 ? 128 : 256, S = v1 ? 256 : 128;
  ok &= run_case({128, 136, 16, 17, L, S, 1, 0, 0, 0}, "M128 N136");
  ok &= run_case({128, 16, 16, 2, L, S, 3, 0, 0, 0}, "M128 N16 x3 accumulate");
  ok &= run_case({64, 72, 8, 9, L, S, 1, 0, 0, 0}, "M64 N72");
  ok &= run_case({64, 8, 8, 1, L, S, 2, 0, 0, 0}, "M64 N8 x2");
  ok &= run_case({64, 16, 8, 2, L, S, 1, 0, 0, 0}, "M64 N16");
  ok &= run_case({64, 24, 8, 3, L, S, 1, 0, 0, 0}, "M64 N24");
  ok &= run_case({64, 40, 8, 5, L, S, 4, 0, 0, 0}, "M64 N40 x4");
  ok &= run_case({64, 72, 8, 9, L, S, 2, 16, 0, 0}, "M64 N72 lane offset 16");
  ok &= run_case({64, 40, 8, 5, L, S, 2, 0, 1, 0}, "M64 N40 bf16 x fp16");
  ok &= run_case({64, 72, 8, 9, L, S, 3, 0, 0, 1}, "M64 N72 stmatrix operands");
  ok &= run_case({128, 136, 16, 17, L, S, 2, 0, 0, 1}, "M128 N136 stmatrix");
  printf(ok ? "ALL OK\n" : "SOME FAILED\n");
  return ok ? 0 : 1;
}
