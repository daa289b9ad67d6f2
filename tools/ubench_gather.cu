// ubench_gather.cu -- which memory path should feed the force sweep's neighbour fetches?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather tools/ubench_gather.cu
//   ./tools/ubench_gather
//
// The force sweep (sph_full.cu) fetches two 16-byte records per neighbour and lane from
// places that differ per lane; ncu showed l1tex__data_pipe_lsu_wavefronts at 86 %.
// This measures, on the access pattern of the sweep (32 consecutive cell-sorted targets,
// each lane picking one of ~52 candidates in one of the 9 neighbour rows):
//   ldg+ldg   : both records through LDG.128 (what k_force_stream did in round 1)
//   tex+tex   : both through the texture path (tex1Dfetch<float4>)
//   ldg+tex   : one each -- is the TEX data pipe additional throughput?
//   lds       : the records staged in shared memory, LDS.128 at lane-dependent slots
//               (conflict-free / random slots), cycles per warp instruction and SM
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int TRIPS = 42;          // neighbour trips per warp of targets
constexpr int ROW = 2445;          // particles per cell row of the 16M scene (256 cells x 9.55)
constexpr int PLANE = ROW * 121;   // particles per z-plane of rows
constexpr int WIN = 52;            // candidates in the union of a warp's runs in one row

__device__ __forceinline__ uint32_t lcg(uint32_t& s)
{
   s = s * 1664525u + 1013904223u;
   return s >> 8;
}

// the sorted index a lane fetches at one trip: warp base + one of 9 rows + window offset
__device__ __forceinline__ int pick(uint32_t& s, int base, int n)
{
   uint32_t r = lcg(s);
   int row = (int)(r % 9u);
   int off = (int)((r >> 4) % (uint32_t)WIN);
   int j = base + (row / 3 - 1) * PLANE + (row % 3 - 1) * ROW + off - WIN / 2;
   j = j < 0 ? j + n : j;
   return j >= n ? j - n : j;
}

template <int MODE>   // 0: ldg+ldg, 1: tex+tex, 2: ldg+tex
__global__ void __launch_bounds__(128, 4) k_gather(const float4* __restrict__ a, const float4* __restrict__ b,
                                                   cudaTextureObject_t ta, cudaTextureObject_t tb, int n,
                                                   float4* __restrict__ out)
{
   const int k = blockIdx.x * blockDim.x + threadIdx.x;
   const int base = (k & ~31) + 16;
   uint32_t s = (uint32_t)k * 2654435761u + 12345u;
   float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 1
   for (int t = 0; t < TRIPS; t += 2)
   {
      int j0 = pick(s, base, n), j1 = pick(s, base, n);
      float4 p0, p1, v0, v1;
      if (MODE == 1) { p0 = tex1Dfetch<float4>(ta, j0); p1 = tex1Dfetch<float4>(ta, j1); }
      else { p0 = __ldg(&a[j0]); p1 = __ldg(&a[j1]); }
      if (MODE == 0) { v0 = __ldg(&b[j0]); v1 = __ldg(&b[j1]); }
      else { v0 = tex1Dfetch<float4>(tb, j0); v1 = tex1Dfetch<float4>(tb, j1); }
      acc.x += p0.x * v0.x + p1.x * v1.x;
      acc.y += p0.y * v0.y + p1.y * v1.y;
      acc.z += p0.z * v0.z + p1.z * v1.z;
      acc.w += p0.w * v0.w + p1.w * v1.w;
   }
   if (k < n)
      out[k] = acc;
}

// a pick closer to the sweep's real pattern (~13 distinct lines per warp-wide load): all lanes are within
// one row of a common row that advances with the trip
__device__ __forceinline__ int pick_near(uint32_t& s, int base, int n, int t)
{
   uint32_t r = lcg(s);
   int row = t * 9 / TRIPS + (int)(r % 3u) - 1;
   row = row < 0 ? 0 : (row > 8 ? 8 : row);
   int off = (int)((r >> 4) % (uint32_t)WIN);
   int j = base + (row / 3 - 1) * PLANE + (row % 3 - 1) * ROW + off - WIN / 2;
   j = j < 0 ? j + n : j;
   return j >= n ? j - n : j;
}

struct __align__(32) Rec32 { float4 p, v; };

// MODE 0: two LDG.128 from two arrays, 1: one 256-bit load from 32-byte records, 2: two LDG.128 from the
// two halves of the 32-byte record (same line)
template <int MODE>
__global__ void __launch_bounds__(128, 4) k_gather_near(const float4* __restrict__ a, const float4* __restrict__ b,
                                                        const Rec32* __restrict__ ab, int n, float4* __restrict__ out)
{
   const int k = blockIdx.x * blockDim.x + threadIdx.x;
   const int base = (k & ~31) + 16;
   uint32_t s = (uint32_t)k * 2654435761u + 12345u;
   float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 1
   for (int t = 0; t < TRIPS; t += 2)
   {
      int j0 = pick_near(s, base, n, t), j1 = pick_near(s, base, n, t + 1);
      float4 p0, p1, v0, v1;
      if (MODE == 0) { p0 = __ldg(&a[j0]); p1 = __ldg(&a[j1]); v0 = __ldg(&b[j0]); v1 = __ldg(&b[j1]); }
      else if (MODE == 1)
      {
         float r0[8], r1[8];
         asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r0[0]), "=f"(r0[1]), "=f"(r0[2]), "=f"(r0[3]), "=f"(r0[4]), "=f"(r0[5]), "=f"(r0[6]), "=f"(r0[7]) : "l"(ab + j0));
         asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r1[0]), "=f"(r1[1]), "=f"(r1[2]), "=f"(r1[3]), "=f"(r1[4]), "=f"(r1[5]), "=f"(r1[6]), "=f"(r1[7]) : "l"(ab + j1));
         p0 = make_float4(r0[0], r0[1], r0[2], r0[3]); v0 = make_float4(r0[4], r0[5], r0[6], r0[7]);
         p1 = make_float4(r1[0], r1[1], r1[2], r1[3]); v1 = make_float4(r1[4], r1[5], r1[6], r1[7]);
      }
      else { p0 = __ldg(&ab[j0].p); v0 = __ldg(&ab[j0].v); p1 = __ldg(&ab[j1].p); v1 = __ldg(&ab[j1].v); }
      acc.x += p0.x * v0.x + p1.x * v1.x;
      acc.y += p0.y * v0.y + p1.y * v1.y;
      acc.z += p0.z * v0.z + p1.z * v1.z;
      acc.w += p0.w * v0.w + p1.w * v1.w;
   }
   if (k < n)
      out[k] = acc;
}

// the same trips with the position record staged in shared memory (LDS.128 at a random slot of a
// 2048-slot window, 32 KB per CTA -> 6 CTAs per SM) and the velocity record through MODE:
//   0: TEX   1: LDG   2: a second LDS.128   3: nothing (LDS only)
template <int MODE>
__global__ void __launch_bounds__(128, 6) k_gather_smem(const float4* __restrict__ a, const float4* __restrict__ b,
                                                        cudaTextureObject_t tb, int n, float4* __restrict__ out)
{
   extern __shared__ float4 tile[];
   const int k = blockIdx.x * blockDim.x + threadIdx.x;
   for (int i = threadIdx.x; i < 2048; i += blockDim.x)
      tile[i] = a[(blockIdx.x * 128 + i) % n];
   __syncthreads();
   const int base = (k & ~31) + 16;
   uint32_t s = (uint32_t)k * 2654435761u + 12345u;
   float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll 1
   for (int t = 0; t < TRIPS; t += 2)
   {
      int j0 = pick(s, base, n), j1 = pick(s, base, n);
      float4 p0 = tile[j0 & 2047], p1 = tile[j1 & 2047], v0, v1;
      if (MODE == 0) { v0 = tex1Dfetch<float4>(tb, j0); v1 = tex1Dfetch<float4>(tb, j1); }
      else if (MODE == 1) { v0 = __ldg(&b[j0]); v1 = __ldg(&b[j1]); }
      else if (MODE == 2) { v0 = tile[(j0 * 7 + 3) & 2047]; v1 = tile[(j1 * 7 + 3) & 2047]; }
      else { v0 = make_float4(1, 1, 1, 1); v1 = v0; }
      acc.x += p0.x * v0.x + p1.x * v1.x;
      acc.y += p0.y * v0.y + p1.y * v1.y;
      acc.z += p0.z * v0.z + p1.z * v1.z;
      acc.w += p0.w * v0.w + p1.w * v1.w;
   }
   if (k < n)
      out[k] = acc;
}

// LDS.128 at lane-dependent slots of a staged tile (SLOTS records of 16 B)
template <int RANDOM>
__global__ void __launch_bounds__(512, 1) k_lds(float4* out, long long* cyc, int slots)
{
   extern __shared__ float4 tile[];
   for (int i = threadIdx.x; i < slots; i += blockDim.x)
      tile[i] = make_float4(i, i + 1, i + 2, i + 3);
   __syncthreads();
   uint32_t s = (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
   float4 acc = make_float4(0, 0, 0, 0);
   const int iters = 2048;
   long long t0 = clock64();
#pragma unroll 1
   for (int it = 0; it < iters; it += 4)
   {
      int j[4];
#pragma unroll
      for (int q = 0; q < 4; q++)
         j[q] = RANDOM ? (int)(lcg(s) % (uint32_t)slots) : (int)((threadIdx.x + 37 * (it + q)) % slots);
#pragma unroll
      for (int q = 0; q < 4; q++)
      {
         float4 v = tile[j[q]];
         acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
   }
   long long t1 = clock64();
   out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
   if (threadIdx.x == 0)
      cyc[blockIdx.x] = t1 - t0;
}

static cudaTextureObject_t make_tex(const float4* p, int n)
{
   cudaResourceDesc rd = {};
   rd.resType = cudaResourceTypeLinear;
   rd.res.linear.devPtr = (void*)p;
   rd.res.linear.desc = cudaCreateChannelDesc<float4>();
   rd.res.linear.sizeInBytes = sizeof(float4) * (size_t)n;
   cudaTextureDesc td = {};
   td.readMode = cudaReadModeElementType;
   cudaTextureObject_t t = 0;
   CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
   return t;
}

int main()
{
   const int n = 16777216;
   float4 *a, *b, *out;
   CK(cudaMalloc(&a, sizeof(float4) * (size_t)n));
   CK(cudaMalloc(&b, sizeof(float4) * (size_t)n));
   CK(cudaMalloc(&out, sizeof(float4) * (size_t)n));
   CK(cudaMemset(a, 0, sizeof(float4) * (size_t)n));
   CK(cudaMemset(b, 0, sizeof(float4) * (size_t)n));
   cudaTextureObject_t ta = make_tex(a, n), tb = make_tex(b, n);
   cudaEvent_t e0, e1;
   CK(cudaEventCreate(&e0));
   CK(cudaEventCreate(&e1));
   const char* names[3] = {"ldg+ldg", "tex+tex", "ldg+tex"};
   for (int rep = 0; rep < 2; rep++)
      for (int mode = 0; mode < 3; mode++)
      {
         CK(cudaEventRecord(e0));
         if (mode == 0) k_gather<0><<<n / 128, 128>>>(a, b, ta, tb, n, out);
         if (mode == 1) k_gather<1><<<n / 128, 128>>>(a, b, ta, tb, n, out);
         if (mode == 2) k_gather<2><<<n / 128, 128>>>(a, b, ta, tb, n, out);
         CK(cudaEventRecord(e1));
         CK(cudaEventSynchronize(e1));
         float ms = 0;
         CK(cudaEventElapsedTime(&ms, e0, e1));
         if (rep)
            printf("gather %-8s : %.3f ms for %d targets x %d trips x 2 records (%.1f clk per warp-trip and SM at 1.965 GHz)\n",
                   names[mode], ms, n, TRIPS, ms * 1e-3 * 1.965e9 * 148.0 / ((double)n / 32 * TRIPS));
      }
   Rec32* ab;
   CK(cudaMalloc(&ab, sizeof(Rec32) * (size_t)n));
   CK(cudaMemset(ab, 0, sizeof(Rec32) * (size_t)n));
   const char* nnames[3] = {"near 2 x ldg128 (2 arrays)", "near 1 x ldg256 (32 B rec)", "near 2 x ldg128 (32 B rec)"};
   for (int rep = 0; rep < 2; rep++)
      for (int mode = 0; mode < 3; mode++)
      {
         CK(cudaEventRecord(e0));
         if (mode == 0) k_gather_near<0><<<n / 128, 128>>>(a, b, ab, n, out);
         if (mode == 1) k_gather_near<1><<<n / 128, 128>>>(a, b, ab, n, out);
         if (mode == 2) k_gather_near<2><<<n / 128, 128>>>(a, b, ab, n, out);
         CK(cudaEventRecord(e1));
         CK(cudaEventSynchronize(e1));
         float ms = 0;
         CK(cudaEventElapsedTime(&ms, e0, e1));
         if (rep)
            printf("gather %-28s : %.3f ms (%.1f clk per warp-trip and SM)\n", nnames[mode], ms,
                   ms * 1e-3 * 1.965e9 * 148.0 / ((double)n / 32 * TRIPS));
      }
   const char* snames[4] = {"lds+tex", "lds+ldg", "lds+lds", "lds only"};
   for (int rep = 0; rep < 2; rep++)
      for (int mode = 0; mode < 4; mode++)
      {
         CK(cudaEventRecord(e0));
         if (mode == 0) k_gather_smem<0><<<n / 128, 128, 32768>>>(a, b, tb, n, out);
         if (mode == 1) k_gather_smem<1><<<n / 128, 128, 32768>>>(a, b, tb, n, out);
         if (mode == 2) k_gather_smem<2><<<n / 128, 128, 32768>>>(a, b, tb, n, out);
         if (mode == 3) k_gather_smem<3><<<n / 128, 128, 32768>>>(a, b, tb, n, out);
         CK(cudaEventRecord(e1));
         CK(cudaEventSynchronize(e1));
         float ms = 0;
         CK(cudaEventElapsedTime(&ms, e0, e1));
         if (rep)
            printf("gather %-8s : %.3f ms (%.1f clk per warp-trip and SM; includes staging 2048 slots per CTA)\n",
                   snames[mode], ms, ms * 1e-3 * 1.965e9 * 148.0 / ((double)n / 32 * TRIPS));
      }
   long long* cyc;
   CK(cudaMalloc(&cyc, sizeof(long long) * 148));
   long long h[148];
   for (int slots : {3000, 6000})
      for (int random = 0; random < 2; random++)
      {
         size_t sm = sizeof(float4) * (size_t)slots;
         CK(cudaFuncSetAttribute(k_lds<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
         CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
         if (random) k_lds<1><<<148, 512, sm>>>(out, cyc, slots); else k_lds<0><<<148, 512, sm>>>(out, cyc, slots);
         CK(cudaDeviceSynchronize());
         CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
         double mean = 0;
         for (int i = 0; i < 148; i++) mean += (double)h[i];
         mean /= 148;
         printf("lds128 %s, %d slots: %.2f clk per warp LDS.128 and SM (16 warps, 2048 loads each)\n",
                random ? "random slots  " : "conflict free ", slots, mean / (16.0 * 2048));
      }
   return 0;
}
