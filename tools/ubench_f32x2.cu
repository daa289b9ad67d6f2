// ubench_f32x2.cu -- issue-rate microbenchmarks behind the density-sweep design:
// is the packed FP32 pipe (FFMA2 / FADD2, sm_100a) twice the scalar FFMA rate per
// issue slot, and what does one density candidate cost in each formulation?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_f32x2 tools/ubench_f32x2.cu
//   ./tools/ubench_f32x2
//
// Prints warp-instructions per clock per SM for dependent-chain-free streams of
// FFMA, FFMA2, FADD2 and an FFMA2 + FMNMX + SHF mix, then cycles per candidate
// (per scheduler) for the scalar and the packed density inner loop over shared
// memory.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
   u64 r;
   asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
   return r;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b)
{
   u64 r;
   asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
   return r;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b)
{
   u64 r;
   asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
   return r;
}
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
   u64 r;
   asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
   return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi)
{
   asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

__global__ void k_ffma(float* out, float a, float b, long long* cyc)
{
   float x[CH];
#pragma unroll
   for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i;
   long long t0 = clock64();
   for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < CH; i++) x[i] = fmaf(x[i], a, b);
   long long t1 = clock64();
   float s = 0;
#pragma unroll
   for (int i = 0; i < CH; i++) s += x[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ffma2(float* out, float a, float b, long long* cyc)
{
   u64 x[CH];
   u64 A = pack2(a, a), B = pack2(b, b);
#pragma unroll
   for (int i = 0; i < CH; i++) x[i] = pack2(threadIdx.x + i, i);
   long long t0 = clock64();
   for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < CH; i++) x[i] = ffma2(x[i], A, B);
   long long t1 = clock64();
   float s = 0;
#pragma unroll
   for (int i = 0; i < CH; i++)
   {
      float lo, hi;
      unpack2(x[i], lo, hi);
      s += lo + hi;
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_fadd2(float* out, float a, float b, long long* cyc)
{
   u64 x[CH];
   u64 B = pack2(b, a);
#pragma unroll
   for (int i = 0; i < CH; i++) x[i] = pack2(threadIdx.x + i, i);
   long long t0 = clock64();
   for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < CH; i++) x[i] = fadd2(x[i], B);
   long long t1 = clock64();
   float s = 0;
#pragma unroll
   for (int i = 0; i < CH; i++)
   {
      float lo, hi;
      unpack2(x[i], lo, hi);
      s += lo + hi;
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FFMA2 + 2 FMNMX + 2 SHF per trip and chain: the instruction mix of the packed sweep
__global__ void k_mix(float* out, float a, float b, long long* cyc)
{
   u64 x[CH];
   unsigned m[CH];
   u64 A = pack2(a, a), B = pack2(b, b);
#pragma unroll
   for (int i = 0; i < CH; i++)
   {
      x[i] = pack2(threadIdx.x + i, i);
      m[i] = i;
   }
   long long t0 = clock64();
   for (int it = 0; it < ITERS; it++)
#pragma unroll
      for (int i = 0; i < CH; i++)
      {
         x[i] = ffma2(x[i], A, B);
         float lo, hi;
         unpack2(x[i], lo, hi);
         m[i] = __funnelshift_l(__float_as_uint(lo), m[i], 1);
         m[i] = __funnelshift_l(__float_as_uint(hi), m[i], 1);
         lo = fminf(lo, 0.5f);
         hi = fminf(hi, 0.25f);
         x[i] = pack2(lo, hi);
      }
   long long t1 = clock64();
   float s = 0;
#pragma unroll
   for (int i = 0; i < CH; i++)
   {
      float lo, hi;
      unpack2(x[i], lo, hi);
      s += lo + hi + m[i];
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// packed and scalar FMAs interleaved (NP packed + NS scalar independent chains): do they
// share one pipe, or does the scalar stream find a second one?
template <int NP, int NS>
__global__ void k_both(float* out, float a, float b, long long* cyc)
{
   u64 x[NP + 1];
   float y[NS + 1];
   u64 A = pack2(a, a), B = pack2(b, b);
#pragma unroll
   for (int i = 0; i < NP; i++) x[i] = pack2(threadIdx.x + i, i);
#pragma unroll
   for (int i = 0; i < NS; i++) y[i] = threadIdx.x - i;
   long long t0 = clock64();
   for (int it = 0; it < ITERS; it++)
   {
#pragma unroll
      for (int i = 0; i < (NP > NS ? NP : NS); i++)
      {
         if (i < NP) x[i] = ffma2(x[i], A, B);
         if (i < NS) y[i] = fmaf(y[i], a, b);
      }
   }
   long long t1 = clock64();
   float s = 0;
#pragma unroll
   for (int i = 0; i < NP; i++)
   {
      float lo, hi;
      unpack2(x[i], lo, hi);
      s += lo + hi;
   }
#pragma unroll
   for (int i = 0; i < NS; i++) s += y[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- density inner loops over shared memory ------------------------------------
constexpr int NC = 2048;   // candidates staged
constexpr int RUN = 256;   // candidates per thread and pass

__global__ void __launch_bounds__(512, 2) k_density_scalar(float* out, unsigned* mout, float hs2, long long* cyc, int reps)
{
   __shared__ float4 sp[NC];
   for (int i = threadIdx.x; i < NC; i += blockDim.x)
      sp[i] = make_float4(0.01f * (i % 97), 0.013f * (i % 89), 0.017f * (i % 83), 1.0f);
   __syncthreads();
   float xi = 0.01f * threadIdx.x, yi = 0.2f, zi = 0.3f;
   const float tmin = -1e-5f * hs2;
   float sum = 0;
   unsigned macc = 0;
   long long t0 = clock64();
   for (int r = 0; r < reps; r++)
   {
      int b = ((threadIdx.x >> 3) * 8 + r * 32) & (NC - RUN - 1);
#pragma unroll 1
      for (int c0 = b; c0 < b + RUN; c0 += 32)
      {
         const float4* p = sp + c0;
         const float4* pe = p + 32;
         unsigned mask = 0;
#pragma unroll 4
         for (; p < pe; p++)
         {
            float4 pj = *p;
            float dx = xi - pj.x, dy = yi - pj.y, dz = zi - pj.z;
            float tt = fmaf(-dz, dz, fmaf(-dy, dy, fmaf(-dx, dx, hs2)));
            mask = __funnelshift_l(__float_as_uint(tmin - tt), mask, 1);
            float tc = fmaxf(tt, 0.0f);
            sum = fmaf(tc * tc, tc, sum);
         }
         macc ^= mask;
      }
   }
   long long t1 = clock64();
   out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
   mout[blockIdx.x * blockDim.x + threadIdx.x] = macc;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// packed: candidates stored as groups of four {x0..x3}{y0..y3}{z0..z3}
__global__ void __launch_bounds__(512, 2) k_density_packed(float* out, unsigned* mout, float hs2, long long* cyc, int reps)
{
   __shared__ float4 sg[NC / 4 * 3];
   for (int g = threadIdx.x; g < NC / 4; g += blockDim.x)
   {
      float4 x, y, z;
      float* xp = &x.x;
      float* yp = &y.x;
      float* zp = &z.x;
      for (int q = 0; q < 4; q++)
      {
         int i = g * 4 + q;
         xp[q] = 0.01f * (i % 97);
         yp[q] = 0.013f * (i % 89);
         zp[q] = 0.017f * (i % 83);
      }
      sg[g * 3] = x;
      sg[g * 3 + 1] = y;
      sg[g * 3 + 2] = z;
   }
   __syncthreads();
   float xi = 0.01f * threadIdx.x, yi = 0.2f, zi = 0.3f;
   const u64 XI = pack2(-xi, -xi), YI = pack2(-yi, -yi), ZI = pack2(-zi, -zi);
   const u64 NH = pack2(-hs2, -hs2);
   const float thr = 1e-5f * hs2;
   const u64 NTHR = pack2(-thr, -thr);
   u64 sum2 = 0;
   unsigned macc = 0;
   long long t0 = clock64();
   for (int r = 0; r < reps; r++)
   {
      int b = ((threadIdx.x >> 3) * 8 + r * 32) & (NC - RUN - 1);
      b &= ~3;
#pragma unroll 1
      for (int c0 = b; c0 < b + RUN; c0 += 32)
      {
         const ulonglong2* p = reinterpret_cast<const ulonglong2*>(sg + (c0 >> 2) * 3);
         unsigned mask = 0;
#pragma unroll 2
         for (int g = 0; g < 8; g++, p += 3)
         {
            ulonglong2 X = p[0], Y = p[1], Z = p[2];
#pragma unroll
            for (int hlf = 0; hlf < 2; hlf++)
            {
               u64 xj = hlf ? X.y : X.x, yj = hlf ? Y.y : Y.x, zj = hlf ? Z.y : Z.x;
               u64 dx = fadd2(xj, XI), dy = fadd2(yj, YI), dz = fadd2(zj, ZI);
               u64 e = ffma2(dx, dx, NH);
               e = ffma2(dy, dy, e);
               e = ffma2(dz, dz, e);            // d2 - hs2
               u64 s = fadd2(e, NTHR);           // sign set <=> inside the enlarged radius
               float s0, s1, e0, e1;
               unpack2(s, s0, s1);
               mask = __funnelshift_l(__float_as_uint(s0), mask, 1);
               mask = __funnelshift_l(__float_as_uint(s1), mask, 1);
               unpack2(e, e0, e1);
               u64 u = pack2(fminf(e0, 0.0f), fminf(e1, 0.0f));
               sum2 = ffma2(fmul2(u, u), u, sum2);
            }
         }
         macc ^= mask;
      }
   }
   long long t1 = clock64();
   float a, b2;
   unpack2(sum2, a, b2);
   out[blockIdx.x * blockDim.x + threadIdx.x] = -(a + b2);
   mout[blockIdx.x * blockDim.x + threadIdx.x] = macc;
   if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double avg_cycles(long long* d_cyc, int blocks)
{
   long long* h = new long long[blocks];
   cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
   double s = 0;
   for (int i = 0; i < blocks; i++) s += (double)h[i];
   delete[] h;
   return s / blocks;
}

int main()
{
   cudaDeviceProp prop;
   cudaGetDeviceProperties(&prop, 0);
   const int sms = prop.multiProcessorCount;
   const int threads = 512, blocks = sms * 2;   // 32 warps per SM
   float* out;
   unsigned* mout;
   long long* cyc;
   cudaMalloc(&out, sizeof(float) * threads * blocks);
   cudaMalloc(&mout, sizeof(unsigned) * threads * blocks);
   cudaMalloc(&cyc, sizeof(long long) * blocks);
   printf("device %s, %d SMs\n", prop.name, sms);
   const double warps_per_sm = 2.0 * threads / 32;
   for (int rep = 0; rep < 2; rep++)
   {
      k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      double c = avg_cycles(cyc, blocks);
      if (rep) printf("FFMA   : %.3f warp-instr/clk/SM\n", warps_per_sm * ITERS * CH / c);
      k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("FFMA2  : %.3f warp-instr/clk/SM (x2 flops)\n", warps_per_sm * ITERS * CH / c);
      k_fadd2<<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("FADD2  : %.3f warp-instr/clk/SM\n", warps_per_sm * ITERS * CH / c);
      k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("MIX    : %.3f trips/clk/SM (1 FFMA2 + 2 FMNMX + 2 SHF per trip)\n", warps_per_sm * ITERS * CH / c);
      k_both<4, 4><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("4 FFMA2 + 4 FFMA : %.3f FMA-lane-equivalents/clk/SM (x32 lanes)\n", warps_per_sm * ITERS * (4 * 2 + 4) / c);
      k_both<4, 2><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("4 FFMA2 + 2 FFMA : %.3f FMA-lane-equivalents/clk/SM\n", warps_per_sm * ITERS * (4 * 2 + 2) / c);
      k_both<6, 2><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("6 FFMA2 + 2 FFMA : %.3f FMA-lane-equivalents/clk/SM\n", warps_per_sm * ITERS * (6 * 2 + 2) / c);
      k_both<8, 0><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("8 FFMA2          : %.3f FMA-lane-equivalents/clk/SM\n", warps_per_sm * ITERS * (8 * 2) / c);
      k_both<0, 8><<<blocks, threads>>>(out, 1.0001f, 0.5f, cyc);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("8 FFMA           : %.3f FMA-lane-equivalents/clk/SM\n", warps_per_sm * ITERS * 8 / c);
      const int reps = 64;
      k_density_scalar<<<blocks, threads>>>(out, mout, 0.01f, cyc, reps);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("density scalar: %.2f cycles per candidate and scheduler\n", c / ((double)reps * RUN * warps_per_sm / 4));
      k_density_packed<<<blocks, threads>>>(out, mout, 0.01f, cyc, reps);
      cudaDeviceSynchronize();
      c = avg_cycles(cyc, blocks);
      if (rep) printf("density packed: %.2f cycles per candidate and scheduler\n", c / ((double)reps * RUN * warps_per_sm / 4));
   }
   cudaError_t e = cudaGetLastError();
   printf("status: %s\n", cudaGetErrorString(e));
   return e != cudaSuccess;
}
