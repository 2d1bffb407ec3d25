// Issue-rate microbenchmarks for the sm_100a SIMT pipes used by rank_pairs (fp32 / packed fp32 / half2 /
// MUFU / shuffle / conversions).  Prints warp-instructions per cycle per SM sub-partition (SMSP).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on one B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

// every kernel: each thread owns CHAINS independent dependency chains, ITERS iterations, fully unrolled x CHAINS
#define KERNEL(name, decl, body, sink)                                                        \
  __global__ void __launch_bounds__(512) name(float* out, int iters) {                        \
    decl;                                                                                     \
    for (int it = 0; it < iters; ++it) {                                                      \
      _Pragma("unroll") for (int c = 0; c < CHAINS; ++c) { body; }                            \
    }                                                                                         \
    sink;                                                                                     \
  }

// scalar fp32 fma, 3 register operands
KERNEL(k_ffma, float x[CHAINS]; float a = out[1]; float b = out[2]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[c]) : "f"(a), "f"(b)),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
// scalar fp32 fma with an immediate
KERNEL(k_ffma_imm, float x[CHAINS]; float a = out[1]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("fma.rn.f32 %0, %0, %1, 0f3F8CCCCD;" : "+f"(x[c]) : "f"(a)),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
KERNEL(k_fadd, float x[CHAINS]; float a = out[1]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[c]) : "f"(a)),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
// packed fp32x2
KERNEL(k_ffma2, unsigned long long x[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1]; unsigned long long b = ((unsigned long long*)out)[2];
       for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[c]) : "l"(a), "l"(b)),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_fadd2, unsigned long long x[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1];
       for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[c]) : "l"(a)),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_fmul2, unsigned long long x[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1];
       for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[c]) : "l"(a)),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
// half2 / bf16x2
KERNEL(k_hfma2, unsigned x[CHAINS]; unsigned a = ((unsigned*)out)[1]; unsigned b = ((unsigned*)out)[2];
       for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b)),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_bfma2, unsigned x[CHAINS]; unsigned a = ((unsigned*)out)[1]; unsigned b = ((unsigned*)out)[2];
       for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b)),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_hmul2, unsigned x[CHAINS]; unsigned a = ((unsigned*)out)[1];
       for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(x[c]) : "r"(a)),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
// MUFU family
KERNEL(k_ex2, float x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3f + c,
       asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c])),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
KERNEL(k_rcp, float x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3f + c + 1,
       asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c])),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
KERNEL(k_tanh, float x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3f + c,
       asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[c])),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
KERNEL(k_ex2_h2, unsigned x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = 0x3c003c00u + threadIdx.x + c,
       asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[c])),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_tanh_h2, unsigned x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = 0x3c003c00u + threadIdx.x + c,
       asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(x[c])),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_tanh_bf2, unsigned x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = 0x3f803f80u + threadIdx.x + c,
       asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(x[c])),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
// conversions
KERNEL(k_cvt_pack_h2, float x[CHAINS]; unsigned y[CHAINS]; for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; y[c] = 0; },
       asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y[c]) : "f"(x[c]), "f"(x[c])); x[c] = __uint_as_float(y[c]),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += y[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_cvt_unpack_h, float x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       unsigned short h = (unsigned short)__float_as_uint(x[c]); asm volatile("cvt.f32.f16 %0, %1;" : "=f"(x[c]) : "h"(h)),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
// shuffle, logic
KERNEL(k_shfl, float x[CHAINS]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       x[c] = __shfl_xor_sync(0xffffffffu, x[c], 4),
       float s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123.f) out[0] = s)
KERNEL(k_lop3, unsigned x[CHAINS]; unsigned a = ((unsigned*)out)[1]; unsigned b = ((unsigned*)out)[2]; for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c,
       asm volatile("lop3.b32 %0, %0, %1, %2, 0xE4;" : "+r"(x[c]) : "r"(a), "r"(b)),
       unsigned s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c]; if (s == 123) out[0] = (float)s)
// mixes: do the pipes overlap?
KERNEL(k_mix_ffma2_ffma, unsigned long long x[CHAINS]; float y[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1]; float fa = out[1];
       for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; y[c] = c; },
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a)); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(y[c]) : "f"(fa)),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c] + (unsigned long long)y[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_mix_ffma2_hfma2, unsigned long long x[CHAINS]; unsigned y[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1]; unsigned ha = ((unsigned*)out)[1];
       for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; y[c] = c; },
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a)); asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(y[c]) : "r"(ha)),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c] + y[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_mix_ffma2_ex2, unsigned long long x[CHAINS]; float y[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1];
       for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; y[c] = c * 0.1f; },
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a));
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a));
       asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y[c])),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c] + (unsigned long long)y[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_mix_ffma2_lop3, unsigned long long x[CHAINS]; unsigned y[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1]; unsigned ha = ((unsigned*)out)[1];
       for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; y[c] = c; },
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a)); asm volatile("lop3.b32 %0, %0, %1, %1, 0xE4;" : "+r"(y[c]) : "r"(ha)),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c] + y[c]; if (s == 123) out[0] = (float)s)
KERNEL(k_mix_ffma2_shfl, unsigned long long x[CHAINS]; float y[CHAINS]; unsigned long long a = ((unsigned long long*)out)[1];
       for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x + c; y[c] = c * 0.1f; },
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a));
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(a));
       y[c] = __shfl_xor_sync(0xffffffffu, y[c], 4),
       unsigned long long s = 0; for (int c = 0; c < CHAINS; ++c) s += x[c] + (unsigned long long)y[c]; if (s == 123) out[0] = (float)s)

// legacy tensor path: mma.sync m16n8k16 f16 -> f32, 4 independent accumulator sets per warp
__global__ void __launch_bounds__(512) k_mma(float* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
  float d[4][4] = {};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int c = 0; c < 4; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 123.f) out[0] = s;
}

template <typename F>
static void run(const char* name, F kern, float* buf, int per_iter, int warps_per_smsp) {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int threads = 128 * warps_per_smsp;   // 4 SMSPs x warps x 32
  const int blocks = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(buf, 64);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(buf, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double inst_per_warp = (double)ITERS * per_iter;
  const double cycles = best * 1e-3 * clk_khz * 1e3;
  // each SMSP runs warps_per_smsp warps
  printf("%-22s warps/SMSP %d : %.3f warp-inst/clk/SMSP  (%.2f clk per inst, %.3f ms, clock attr %d MHz)\n", name, warps_per_smsp,
         inst_per_warp * warps_per_smsp / cycles, cycles / (inst_per_warp * warps_per_smsp), best, clk_khz / 1000);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(err));
}

int main() {
  float* buf;
  cudaMalloc(&buf, 1 << 20);
  cudaMemset(buf, 0, 1 << 20);
#define RUN(k, n) for (int w : {1, 2, 4}) run(#k, k, buf, n, w)
  RUN(k_ffma, CHAINS);
  RUN(k_ffma_imm, CHAINS);
  RUN(k_fadd, CHAINS);
  RUN(k_ffma2, CHAINS);
  RUN(k_fadd2, CHAINS);
  RUN(k_fmul2, CHAINS);
  RUN(k_hfma2, CHAINS);
  RUN(k_bfma2, CHAINS);
  RUN(k_hmul2, CHAINS);
  RUN(k_ex2, CHAINS);
  RUN(k_rcp, CHAINS);
  RUN(k_tanh, CHAINS);
  RUN(k_ex2_h2, CHAINS);
  RUN(k_tanh_h2, CHAINS);
  RUN(k_tanh_bf2, CHAINS);
  RUN(k_cvt_pack_h2, CHAINS);
  RUN(k_cvt_unpack_h, CHAINS);
  RUN(k_shfl, CHAINS);
  RUN(k_lop3, CHAINS);
  RUN(k_mix_ffma2_ffma, 2 * CHAINS);
  RUN(k_mix_ffma2_hfma2, 2 * CHAINS);
  RUN(k_mix_ffma2_ex2, 5 * CHAINS);
  RUN(k_mix_ffma2_lop3, 2 * CHAINS);
  RUN(k_mix_ffma2_shfl, 5 * CHAINS);
  RUN(k_mma, 4);
  return 0;
}
