// sph_capi.cu -- the C ABI of include/sphb200.h: context lifetime, parameter
// derivation (the reference constructor's own expressions), host<->HBM state
// transfer, step dispatch and per-step scalars.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <limits>
#include <new>

#include "sph_internal.h"

int sph_full_unsort(sphb200_ctx* ctx);
int sph_full_tile_count(const sphb200_ctx* ctx);

namespace
{

thread_local std::string g_error;   // errors raised without a context

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_pack_state(int n, const float* __restrict__ pos_xyz,
                                                          const float* __restrict__ vel_xyz,
                                                          const float* __restrict__ mass, float4* __restrict__ pos4,
                                                          float4* __restrict__ vel4, int* __restrict__ mass_differs)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   float m = mass ? mass[i] : 1.0f;
   if (mass && m != mass[0])
      *mass_differs = 1;      // benign race: every writer stores the same value
   pos4[i] = make_float4(pos_xyz[3 * (size_t)i], pos_xyz[3 * (size_t)i + 1], pos_xyz[3 * (size_t)i + 2], m);
   vel4[i] = make_float4(vel_xyz[3 * (size_t)i], vel_xyz[3 * (size_t)i + 1], vel_xyz[3 * (size_t)i + 2], 0.0f);
}

// the two halves of k_pack_state for sphb200_step_host (positions + masses first,
// velocities while the step is already running)
__global__ void __launch_bounds__(kThreads) k_pack_pos(int n, const float* __restrict__ pos_xyz,
                                                        const float* __restrict__ mass, float4* __restrict__ pos4,
                                                        int* __restrict__ mass_differs)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   float m = mass ? mass[i] : 1.0f;
   if (mass && m != mass[0])
      *mass_differs = 1;
   pos4[i] = make_float4(pos_xyz[3 * (size_t)i], pos_xyz[3 * (size_t)i + 1], pos_xyz[3 * (size_t)i + 2], m);
}

__global__ void __launch_bounds__(kThreads) k_pack_vel(int n, const float* __restrict__ vel_xyz,
                                                        float4* __restrict__ vel4)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   vel4[i] = make_float4(vel_xyz[3 * (size_t)i], vel_xyz[3 * (size_t)i + 1], vel_xyz[3 * (size_t)i + 2], 0.0f);
}

// float4 -> xyz-interleaved (Particle::mPosition layout, particle.h:15)
__global__ void __launch_bounds__(kThreads) k_unpack_xyz(int n, const float4* __restrict__ src,
                                                          float* __restrict__ dst_xyz)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= n)
      return;
   float4 v = src[i];
   dst_xyz[3 * (size_t)i] = v.x;
   dst_xyz[3 * (size_t)i + 1] = v.y;
   dst_xyz[3 * (size_t)i + 2] = v.z;
}

__global__ void __launch_bounds__(kThreads) k_unpack_w(int n, const float4* __restrict__ src, float* __restrict__ dst)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n)
      dst[i] = src[i].w;
}

int blocks_for(int n) { return (n + kThreads - 1) / kThreads; }

// SPH::SPH() (sph.cpp:47-95): same expressions, same types, same rounding.
void derive(const SphParams& p, SphDerived& d)
{
   float h = p.h;
   float scale = p.simulation_scale;
   d.h2 = (float)pow((double)h, 2.0);
   d.h_times2 = h * 2.0f;
   d.h_times2_inv = 1.0f / d.h_times2;
   d.h_scaled = h * scale;
   d.h_scaled2 = (float)pow((double)(h * scale), 2.0);
   d.h_scaled6 = (float)pow((double)(h * scale), 6.0);
   d.h_scaled9 = (float)pow((double)(h * scale), 9.0);
   d.grid_cell_count = p.grid_x * p.grid_y * p.grid_z;
   d.cell_size = 2.0f * h;
   d.max_x = d.cell_size * (float)p.grid_x;
   d.max_y = d.cell_size * (float)p.grid_y;
   d.max_z = d.cell_size * (float)p.grid_z;
   d.total_steps = (int)round(1.0f / p.time_step);
   if (p.central_pos[0] < 0.0f && p.central_pos[1] < 0.0f && p.central_pos[2] < 0.0f)
   {
      d.central_pos[0] = d.max_x * 0.5f;
      d.central_pos[1] = d.max_y * 0.5f;
      d.central_pos[2] = d.max_z * 0.5f;
   }
   else
      for (int k = 0; k < 3; k++)
         d.central_pos[k] = p.central_pos[k];
   d.softening = p.softening < 0.0f ? d.h_scaled : p.softening;
   d.cfl_limit2 = p.cfl_limit * p.cfl_limit;
   d.kernel1 = 315.0f / (64.0f * (float)(M_PI) * d.h_scaled9);
   d.kernel2 = -45.0f / ((float)(M_PI) * d.h_scaled6);
   d.kernel3 = -d.kernel2;
}

int validate(const SphParams& p, std::string& why)
{
   if (p.particle_count < 0) { why = "particle_count < 0"; return 1; }
   if (p.grid_x < 1 || p.grid_y < 1 || p.grid_z < 1) { why = "grid cells must be >= 1"; return 1; }
   if ((long long)p.grid_x * p.grid_y * p.grid_z * 8ll >= (1ll << 31)) { why = "grid too large for 32-bit cell keys"; return 1; }
   if (p.examine_count < 9) { why = "examine_count must be >= 9 (K=8 windows, sph.cpp:679)"; return 1; }
   if (!(p.h > 0.0f)) { why = "h must be > 0"; return 1; }
   if (!(p.simulation_scale > 0.0f)) { why = "simulation_scale must be > 0"; return 1; }
   if (p.neighbor_mode != SPHB200_NEIGHBORS_REFERENCE_SAMPLED && p.neighbor_mode != SPHB200_NEIGHBORS_FULL)
   {
      why = "unknown neighbor_mode";
      return 1;
   }
   if ((long long)p.particle_count * p.examine_count >= (1ll << 40)) { why = "neighbour table too large"; return 1; }
   return 0;
}

template <typename T>
cudaError_t dev_alloc(T** p, size_t count)
{
   return cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
}

int alloc_lists(sphb200_ctx* ctx)
{
   if (ctx->nbr_idx)
      return SPHB200_OK;
   size_t cap = (size_t)ctx->capacity * (size_t)ctx->params.examine_count;
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->nbr_idx, cap));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->nbr_dist, cap));
   return SPHB200_OK;
}

int fetch_scalars(sphb200_ctx* ctx)
{
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(&ctx->h_scalars, ctx->d_scalars, sizeof(StepScalars), cudaMemcpyDeviceToHost,
                                       ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   return SPHB200_OK;
}

void snapshots_free(sphb200_ctx* ctx)
{
   sphb200_ctx::Snapshots* sn = ctx->snap;
   if (!sn)
      return;
   if (sn->stream)
   {
      cudaStreamSynchronize(sn->stream);
      cudaStreamDestroy(sn->stream);
   }
   if (sn->staged) cudaEventDestroy(sn->staged);
   for (int b = 0; b < 2; b++)
   {
      if (sn->done[b]) cudaEventDestroy(sn->done[b]);
      if (sn->host_pos[b]) cudaFreeHost(sn->host_pos[b]);
      if (sn->host_cnt[b]) cudaFreeHost(sn->host_cnt[b]);
   }
   if (sn->dev_pos) cudaFree(sn->dev_pos);
   if (sn->dev_cnt) cudaFree(sn->dev_cnt);
   delete sn;
   ctx->snap = nullptr;
}

int snapshots_init(sphb200_ctx* ctx)
{
   if (ctx->snap && ctx->snap->ready)
      return SPHB200_OK;
   sphb200_ctx::Snapshots* sn = ctx->snap ? ctx->snap : new (std::nothrow) sphb200_ctx::Snapshots();
   if (!sn)
      return sph_fail(ctx, SPHB200_E_INVALID, "out of host memory");
   ctx->snap = sn;
   const size_t n = (size_t)(ctx->capacity > 0 ? ctx->capacity : 1), cells = (size_t)ctx->cells_voxel;
   SPH_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&sn->stream, cudaStreamNonBlocking));
   SPH_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&sn->staged, cudaEventDisableTiming));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&sn->dev_pos, sizeof(float) * 3 * n));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&sn->dev_cnt, sizeof(uint32_t) * cells));
   for (int b = 0; b < 2; b++)
   {
      SPH_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&sn->done[b], cudaEventDisableTiming));
      SPH_CUDA_CHECK(ctx, cudaMallocHost((void**)&sn->host_pos[b], sizeof(float) * 3 * n));
      SPH_CUDA_CHECK(ctx, cudaMallocHost((void**)&sn->host_cnt[b], sizeof(int) * cells));
      sn->step[b] = -1;
      sn->what[b] = 0;
      sn->in_flight[b] = false;
   }
   sn->newest = -1;
   sn->ready = true;
   return SPHB200_OK;
}

}  // namespace

int sph_fail(sphb200_ctx* ctx, int code, const std::string& msg)
{
   if (ctx)
      ctx->error = msg;
   else
      g_error = msg;
   return code;
}

DevParams sph_dev_params(const sphb200_ctx* ctx)
{
   const SphParams& p = ctx->params;
   const SphDerived& d = ctx->derived;
   DevParams P;
   memset(&P, 0, sizeof(P));
   P.n = ctx->n_local;
   P.n_owned = ctx->n_owned;
   P.gx = p.grid_x; P.gy = p.grid_y; P.gz = p.grid_z;
   P.fx = 2 * p.grid_x; P.fy = 2 * p.grid_y; P.fz = 2 * p.grid_z;
   P.examine = p.examine_count;
   P.use_gravity = p.use_uniform_gravity;
   P.use_walls = p.use_wall_collision;
   P.h = p.h; P.h2 = d.h2; P.h_times2 = d.h_times2; P.h_times2_inv = d.h_times2_inv;
   P.hs = d.h_scaled; P.hs2 = d.h_scaled2;
   P.k1 = d.kernel1; P.k2 = d.kernel2; P.k3 = d.kernel3;
   P.scale = p.simulation_scale;
   P.dt = p.time_step;
   P.pos_dt = p.time_step * (1.0f / p.simulation_scale);   // sph.cpp:49, 956
   P.rho0 = p.rho0; P.stiffness = p.stiffness; P.viscosity = p.viscosity; P.damping = p.damping;
   P.cfl = p.cfl_limit; P.cfl2 = d.cfl_limit2;
   P.neg_gm = -p.grav_constant * p.central_mass;
   P.gm = p.grav_constant * p.central_mass;
   P.cx = d.central_pos[0]; P.cy = d.central_pos[1]; P.cz = d.central_pos[2];
   P.softening = d.softening;
   P.gvx = p.gravity[0]; P.gvy = p.gravity[1]; P.gvz = p.gravity[2];
   P.max_x = d.max_x; P.max_y = d.max_y; P.max_z = d.max_z;
   P.defer_velocity = ctx->deferred_vel_event != nullptr;
   if (ctx->comm)
      sph_comm_dev_params(ctx, P);
   return P;
}

// The step is a fixed sequence of 9-11 launches with fixed arguments (everything a kernel
// needs lives in device memory or in the by-value DevParams), so it is captured once
// into a CUDA graph and replayed (at the reference's default 32 768 particles: 254 ->
// 236 us per step; the rest is the serial neighbour walk of k_find_sampled).  The
// graph is rebuilt after anything that changes an argument (set_params, upload_state,
// set_stream).  Slab contexts (NCCL exchange per step) and timed steps (events between
// the kernels) launch directly.
void sph_graph_invalidate(sphb200_ctx* ctx)
{
   if (ctx->graph_exec)
      cudaGraphExecDestroy(ctx->graph_exec);
   ctx->graph_exec = nullptr;
}

static int step_once(sphb200_ctx* ctx)
{
   return ctx->params.neighbor_mode == SPHB200_NEIGHBORS_FULL ? sph_step_full(ctx) : sph_step_sampled(ctx);
}

static int step_graph(sphb200_ctx* ctx, int n_steps)
{
   cudaStream_t st = ctx->stream;
   if (!ctx->graph_exec)
   {
      const long long l0 = ctx->launches;
      SPH_CUDA_CHECK(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      int rc = step_once(ctx);
      cudaGraph_t g = nullptr;
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (rc || e != cudaSuccess)
      {
         if (g)
            cudaGraphDestroy(g);
         cudaGetLastError();
         return rc ? rc : sph_fail(ctx, SPHB200_E_CUDA, std::string("step capture: ") + cudaGetErrorString(e));
      }
      e = cudaGraphInstantiate(&ctx->graph_exec, g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess)
      {
         ctx->graph_exec = nullptr;
         return sph_fail(ctx, SPHB200_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
      }
      ctx->graph_launches = ctx->launches - l0;
      ctx->launches = l0;
   }
   for (int s = 0; s < n_steps; s++)
   {
      SPH_CUDA_CHECK(ctx, cudaGraphLaunch(ctx->graph_exec, st));
      ctx->launches += ctx->graph_launches;
      ctx->steps_done++;
   }
   // host-side view of what a step leaves behind (the capture ran the same code once)
   const bool full = ctx->params.neighbor_mode == SPHB200_NEIGHBORS_FULL;
   ctx->lists_valid = !full;
   ctx->snapshot_valid = full;
   ctx->stream_valid = full && ctx->params.kernel_variant != 1;
   ctx->unsorted_valid = false;
   ctx->voxel_ids_valid = true;
   ctx->idx_order = ctx->idx_sorted;
   ctx->stepped = true;
   return SPHB200_OK;
}

extern "C" {

int sphb200_default_params(SphParams* p)
{
   if (!p)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null params");
   memset(p, 0, sizeof(*p));
   p->particle_count = 32 * 1024;          // M * 1024, M = 32 (sph.cpp:29-31, 59)
   p->grid_x = p->grid_y = p->grid_z = 32; // sph.cpp:60-62
   p->examine_count = 32;                  // sph.cpp:98
   p->neighbor_mode = SPHB200_NEIGHBORS_REFERENCE_SAMPLED;
   p->h = 0.1f;                            // sph.cpp:47
   p->simulation_scale = 1.0f;             // sph.cpp:48
   p->time_step = 0.001f;                  // sph.cpp:70
   p->rho0 = 0.1f;                         // sph.cpp:74
   p->stiffness = 0.001f;                  // sph.cpp:75
   p->viscosity = 0.01f;                   // sph.cpp:77
   p->damping = 0.001f;                    // sph.cpp:78
   p->cfl_limit = 10000.0f;                // sph.cpp:89
   p->grav_constant = 4.3009e-3f;          // sph.cpp:80
   p->central_mass = 1e+5f;                // sph.cpp:81
   p->central_pos[0] = p->central_pos[1] = p->central_pos[2] = -1.0f;   // box centre (sph.cpp:83-85)
   p->softening = -1.0f;                   // h * scale (sph.cpp:86)
   return SPHB200_OK;
}

int sphb200_derive(const SphParams* p, SphDerived* out)
{
   if (!p || !out)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null argument");
   std::string why;
   if (validate(*p, why))
      return sph_fail(nullptr, SPHB200_E_INVALID, why);
   derive(*p, *out);
   return SPHB200_OK;
}

const char* sphb200_last_error(const sphb200_ctx* ctx)
{
   return ctx ? ctx->error.c_str() : g_error.c_str();
}

int sphb200_create(const SphParams* p, int device, sphb200_ctx** out)
{
   if (!p || !out)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null argument");
   *out = nullptr;
   std::string why;
   if (validate(*p, why))
      return sph_fail(nullptr, SPHB200_E_INVALID, why);
   int ndev = 0;
   cudaError_t e = cudaGetDeviceCount(&ndev);
   if (e != cudaSuccess || ndev == 0)
      return sph_fail(nullptr, SPHB200_E_CUDA,
                      std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
   if (device < 0)
   {
      if (cudaGetDevice(&device) != cudaSuccess)
         device = 0;
   }
   if (device >= ndev)
      return sph_fail(nullptr, SPHB200_E_INVALID, "device index out of range");
   sphb200_ctx* ctx = new (std::nothrow) sphb200_ctx();
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "out of host memory");
   ctx->params = *p;
   derive(*p, ctx->derived);
   ctx->device = device;
   ctx->capacity = p->particle_count;
   ctx->n_local = p->particle_count;
   ctx->n_owned = p->particle_count;
   const char* no_graph = getenv("SPHB200_NO_GRAPH");
   ctx->use_graph = !(no_graph && no_graph[0] == '1');
   ctx->cells_voxel = p->grid_x * p->grid_y * p->grid_z;
   ctx->cells_fine = 8 * ctx->cells_voxel;
   ctx->cells_alloc = p->neighbor_mode == SPHB200_NEIGHBORS_FULL ? ctx->cells_fine : ctx->cells_voxel;
   *out = ctx;   // from here on errors are reported on the context; caller destroys it
   SPH_CUDA_CHECK(ctx, cudaSetDevice(device));
   SPH_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
   ctx->own_stream = true;
   const size_t n = (size_t)ctx->capacity;
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->pos4, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->vel4, n));
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->pos4, 0, sizeof(float4) * (n ? n : 1), ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->vel4, 0, sizeof(float4) * (n ? n : 1), ctx->stream));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->keys, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->keys_sorted, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->idx_iota, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->idx_sorted, n));
   // the two force-sweep record arrays are ONE allocation (s_velB4 = s_posA4 + n rounded up to 32): the AoS build of the
   // sweeps (SPH_FORCE_AOS) lays 32-byte records over it; also the upload / download staging
   const size_t n_rec = (n + 31) & ~(size_t)31;   // 512-byte steps: the second array is also bound as a linear texture
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->s_posA4, 2 * n_rec));
   ctx->s_velB4 = ctx->s_posA4 + n_rec;
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->stage_f, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->nbr_count, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->rho, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->acc4, n));
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->voxel_id, n));
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->rho, 0, sizeof(float) * (n ? n : 1), ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->acc4, 0, sizeof(float4) * (n ? n : 1), ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->nbr_count, 0, sizeof(int) * (n ? n : 1), ctx->stream));
   int max_blocks = (int)((n + 127) / 128) + 1;
   if (p->neighbor_mode == SPHB200_NEIGHBORS_FULL)
   {
      SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->s_pos4, n));
      SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->s_rho, n));
      SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->s_acc4, n));
      SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->s_count, n));
      int tiles = sph_full_tile_count(ctx);
      if (tiles > max_blocks)
         max_blocks = tiles;
      int rc = sph_full_configure(ctx);
      if (rc)
         return rc;
   }
   else
   {
      int rc = alloc_lists(ctx);
      if (rc)
         return rc;
   }
   ctx->partial_blocks = max_blocks;
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->d_block_partials, 2 * (size_t)max_blocks + 2 * 256)   /* + second-stage partials */);
   SPH_CUDA_CHECK(ctx, dev_alloc(&ctx->d_scalars, 1));
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(StepScalars), ctx->stream));
   for (int i = 0; i < 8; i++)
      SPH_CUDA_CHECK(ctx, cudaEventCreate(&ctx->ev[i]));
   int rc = sph_grid_setup(ctx);
   if (rc)
      return rc;
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   return SPHB200_OK;
}

int sphb200_destroy(sphb200_ctx* ctx)
{
   if (!ctx)
      return SPHB200_OK;
   cudaSetDevice(ctx->device);
   if (ctx->stream)
      cudaStreamSynchronize(ctx->stream);
   sph_comm_free(ctx);
   sph_graph_invalidate(ctx);
   if (ctx->upload_stream)
   {
      cudaStreamDestroy(ctx->upload_stream);
      cudaEventDestroy(ctx->upload_ev[0]);
      cudaEventDestroy(ctx->upload_ev[1]);
   }
   snapshots_free(ctx);
   if (ctx->h_report) cudaFreeHost(ctx->h_report);
   if (ctx->tex_posA) cudaDestroyTextureObject(ctx->tex_posA);
   if (ctx->tex_velB) cudaDestroyTextureObject(ctx->tex_velB);
   sph_grid_free(ctx);
   void* bufs[] = {ctx->pos4, ctx->vel4, ctx->gid, ctx->keys, ctx->keys_sorted, ctx->idx_iota, ctx->idx_sorted,
                   ctx->slot_state, ctx->idx_fixed, ctx->s_pos4, ctx->s_posA4, ctx->s_rho,
                   ctx->s_acc4, ctx->s_count, ctx->nbr_idx, ctx->nbr_dist, ctx->nbr_count, ctx->rho, ctx->acc4,
                   ctx->voxel_id, ctx->vg_count, ctx->vg_start, ctx->vg_members, ctx->vg_keys,
                   ctx->d_scalars, ctx->d_block_partials, ctx->stage_f, ctx->hit_rec, ctx->hit_info, ctx->tile_list,
                   ctx->tile_ctl};
   for (void* b : bufs)
      if (b)
         cudaFree(b);
   for (int i = 0; i < 8; i++)
      if (ctx->ev[i])
         cudaEventDestroy(ctx->ev[i]);
   if (ctx->own_stream && ctx->stream)
      cudaStreamDestroy(ctx->stream);
   delete ctx;
   return SPHB200_OK;
}

int sphb200_get_params(const sphb200_ctx* ctx, SphParams* out)
{
   if (!ctx || !out)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null argument");
   *out = ctx->params;
   return SPHB200_OK;
}

int sphb200_get_derived(const sphb200_ctx* ctx, SphDerived* out)
{
   if (!ctx || !out)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null argument");
   *out = ctx->derived;
   return SPHB200_OK;
}

// runtime scalars only -- the reference's setters (sph.cpp:1225-1289) plus the
// constants a harness pokes through protected members.  Structural fields
// (counts, grid, h, scale, mode) are fixed at create.
int sphb200_set_params(sphb200_ctx* ctx, const SphParams* p)
{
   if (!ctx || !p)
      return sph_fail(ctx, SPHB200_E_INVALID, "null argument");
   const SphParams& c = ctx->params;
   if (p->particle_count != c.particle_count || p->grid_x != c.grid_x || p->grid_y != c.grid_y ||
       p->grid_z != c.grid_z || p->examine_count != c.examine_count || p->neighbor_mode != c.neighbor_mode ||
       p->h != c.h || p->simulation_scale != c.simulation_scale)
      return sph_fail(ctx, SPHB200_E_INVALID,
                      "set_params: particle_count/grid/examine_count/neighbor_mode/h/scale are fixed at create");
   ctx->params = *p;
   derive(ctx->params, ctx->derived);
   sph_graph_invalidate(ctx);
   return SPHB200_OK;
}

int sphb200_set_stream(sphb200_ctx* ctx, void* cuda_stream)
{
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null context");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   sph_graph_invalidate(ctx);
   if (ctx->own_stream)
      cudaStreamDestroy(ctx->stream);
   if (cuda_stream)
   {
      ctx->stream = (cudaStream_t)cuda_stream;
      ctx->own_stream = false;
   }
   else
   {
      SPH_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
   }
   return SPHB200_OK;
}

int sphb200_upload_state(sphb200_ctx* ctx, const float* pos_xyz, const float* vel_xyz, const float* mass)
{
   if (!ctx || !pos_xyz || !vel_xyz)
      return sph_fail(ctx, SPHB200_E_INVALID, "upload_state: null argument");
   if (ctx->comm)
      return sph_fail(ctx, SPHB200_E_INVALID, "upload_state: slab contexts use sphb200_upload_slab");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   const int n = ctx->capacity;
   cudaStream_t st = ctx->stream;
   if (n > 0)
   {
      float* d_pos = reinterpret_cast<float*>(ctx->s_posA4);
      float* d_vel = reinterpret_cast<float*>(ctx->s_velB4);
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pos, pos_xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_vel, vel_xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
      if (mass)
         SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->stage_f, mass, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
      int* d_flag = &ctx->d_scalars->overflow;   // free between steps (reset at the start of every step)
      SPH_CUDA_CHECK(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
      k_pack_state<<<blocks_for(n), kThreads, 0, st>>>(n, d_pos, d_vel, mass ? ctx->stage_f : nullptr, ctx->pos4,
                                                       ctx->vel4, d_flag);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
      // all masses equal (the reference never sets anything but 1, sph.cpp:88): lets the
      // density sweep factor the mass out of its inner loop.  Checked on the device while
      // packing (a host loop over 16.7M masses cost more than the whole step).
      int differs = 0;
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(&differs, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
      ctx->uniform_mass = differs == 0;
   }
   else
      ctx->uniform_mass = true;
   ctx->n_local = ctx->n_owned = n;
   sph_graph_invalidate(ctx);
   ctx->lists_valid = false;
   ctx->snapshot_valid = false;
   ctx->stream_valid = false;
   ctx->voxel_ids_valid = false;
   ctx->unsorted_valid = false;
   ctx->stepped = false;
   return SPHB200_OK;
}

int sphb200_step(sphb200_ctx* ctx, int n_steps)
{
   if (!ctx || n_steps < 0)
      return sph_fail(ctx, SPHB200_E_INVALID, "step: bad argument");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   if (ctx->use_graph && !ctx->comm && !ctx->params.enable_timers && n_steps > 0 && ctx->n_local > 0)
      return step_graph(ctx, n_steps);
   for (int s = 0; s < n_steps; s++)
   {
      int rc;
      if (ctx->comm)
      {
         rc = sph_comm_exchange(ctx);
         if (rc)
            return rc;
      }
      rc = ctx->params.neighbor_mode == SPHB200_NEIGHBORS_FULL ? sph_step_full(ctx) : sph_step_sampled(ctx);
      if (rc)
         return rc;
      ctx->stepped = true;
      ctx->steps_done++;
   }
   return SPHB200_OK;
}

int sphb200_synchronize(sphb200_ctx* ctx)
{
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null context");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   return sph_comm_check(ctx);
}

int sphb200_download(sphb200_ctx* ctx, int field, void* dst, size_t dst_bytes)
{
   if (!ctx || !dst)
      return sph_fail(ctx, SPHB200_E_INVALID, "download: null argument");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   const int n = ctx->n_local;
   const size_t E = (size_t)ctx->params.examine_count;
   cudaStream_t st = ctx->stream;
   const bool full = ctx->params.neighbor_mode == SPHB200_NEIGHBORS_FULL;
   size_t need = 0;
   const void* src = nullptr;
   // Slab contexts index slots, not particles: FREE slots have no voxel, the voxel ids are
   // global while the local tables cover one slab, and slot order is not particle order.
   // The per-particle views of a slab go through sphb200_download_slab; the voxel-grid views
   // of the reference (mVoxelIds / mGrid, sph.h:142-146) exist on single-GPU contexts only.
   if (ctx->comm && (field == SPHB200_F_VOXEL_ID || field == SPHB200_F_VOXEL_COORD || field == SPHB200_F_GRID_START ||
                     field == SPHB200_F_GRID_MEMBERS || field == SPHB200_F_CELL_COUNT || field == SPHB200_F_FINE_KEY))
      return sph_fail(ctx, SPHB200_E_INVALID, "download: voxel / grid views are not defined on a slab context");
   switch (field)
   {
   case SPHB200_F_POSITION:
   case SPHB200_F_VELOCITY:
   case SPHB200_F_ACCELERATION:
   {
      need = sizeof(float) * 3 * (size_t)n;
      const float4* from = field == SPHB200_F_POSITION ? ctx->pos4 : field == SPHB200_F_VELOCITY ? ctx->vel4 : ctx->acc4;
      if (field == SPHB200_F_ACCELERATION && full && !ctx->unsorted_valid)
      {
         int rc = sph_full_unsort(ctx);
         if (rc)
            return rc;
      }
      float* stage = reinterpret_cast<float*>(field == SPHB200_F_VELOCITY ? ctx->s_velB4 : ctx->s_posA4);
      if (n > 0)
      {
         k_unpack_xyz<<<blocks_for(n), kThreads, 0, st>>>(n, from, stage);
         ctx->launches++;
         SPH_CUDA_CHECK(ctx, cudaGetLastError());
      }
      src = stage;
      break;
   }
   case SPHB200_F_MASS:
      need = sizeof(float) * (size_t)n;
      if (n > 0)
      {
         k_unpack_w<<<blocks_for(n), kThreads, 0, st>>>(n, ctx->pos4, ctx->stage_f);
         ctx->launches++;
         SPH_CUDA_CHECK(ctx, cudaGetLastError());
      }
      src = ctx->stage_f;
      break;
   case SPHB200_F_DENSITY:
   case SPHB200_F_NEIGHBOR_COUNT:
      if (full && !ctx->unsorted_valid)
      {
         int rc = sph_full_unsort(ctx);
         if (rc)
            return rc;
      }
      need = (field == SPHB200_F_DENSITY ? sizeof(float) : sizeof(int)) * (size_t)n;
      src = field == SPHB200_F_DENSITY ? (const void*)ctx->rho : (const void*)ctx->nbr_count;
      break;
   case SPHB200_F_VOXEL_ID:
   {
      int rc = sph_refresh_voxel_ids(ctx);
      if (rc)
         return rc;
      need = sizeof(int) * (size_t)n;
      src = ctx->voxel_id;
      break;
   }
   case SPHB200_F_VOXEL_COORD:
   {
      // mVoxelCoords (sph.h:144): decoded on the host from the voxel ids
      need = sizeof(int) * 3 * (size_t)n;
      if (dst_bytes < need)
         return sph_fail(ctx, SPHB200_E_INVALID, "download: destination too small");
      int rc = sph_refresh_voxel_ids(ctx);
      if (rc)
         return rc;
      int* out = static_cast<int*>(dst);
      int* ids = out + 2 * (size_t)n;   // tail of the destination as scratch
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(ids, ctx->voxel_id, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
      SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
      const int gx = ctx->params.grid_x, gy = ctx->params.grid_y;
      for (int i = 0; i < n; i++)
      {
         int id = ids[i];
         out[3 * (size_t)i] = id % gx;
         out[3 * (size_t)i + 1] = (id / gx) % gy;
         out[3 * (size_t)i + 2] = id / (gx * gy);
      }
      return SPHB200_OK;
   }
   case SPHB200_F_GRID_START:
   case SPHB200_F_GRID_MEMBERS:
   case SPHB200_F_CELL_COUNT:
      return sph_download_grid(ctx, field, dst, dst_bytes);
   case SPHB200_F_NEIGHBOR_INDEX:
   case SPHB200_F_NEIGHBOR_DISTANCE:
      if (!ctx->lists_valid)
         return sph_fail(ctx, SPHB200_E_INVALID,
                         full ? "download: call sphb200_build_neighbor_lists after a FULL-mode step first"
                              : "download: no step has produced neighbour lists yet");
      need = sizeof(uint32_t) * (size_t)n * E;
      src = field == SPHB200_F_NEIGHBOR_INDEX ? (const void*)ctx->nbr_idx : (const void*)ctx->nbr_dist;
      break;
   case SPHB200_F_FINE_KEY:
      if (!full || !ctx->snapshot_valid)
         return sph_fail(ctx, SPHB200_E_INVALID, "download: fine keys exist after a FULL-mode step only");
      need = sizeof(int) * (size_t)n;
      src = ctx->keys;
      break;
   default:
      return sph_fail(ctx, SPHB200_E_INVALID, "download: unknown field");
   }
   if (dst_bytes < need)
      return sph_fail(ctx, SPHB200_E_INVALID, "download: destination too small");
   if (need)
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, st));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
   return SPHB200_OK;
}

// Upload for sphb200_step_host in FULL mode: binning and the density sweep only need the
// positions (and masses), so those go first and the step starts on them while the
// velocities are still crossing PCIe on a second stream; the step joins that stream
// before the force sweep (sph_step_full: deferred_vel_event, k_gather_vel).
static int step_host_overlapped(sphb200_ctx* ctx, const float* pos_xyz, const float* vel_xyz, const float* mass)
{
   const int n = ctx->capacity;
   cudaStream_t st = ctx->stream;
   if (!ctx->upload_stream)
   {
      SPH_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; i++)
         SPH_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->upload_ev[i], cudaEventDisableTiming));
   }
   cudaStream_t su = ctx->upload_stream;
   float* d_pos = reinterpret_cast<float*>(ctx->s_posA4);   // staging; free until the density sweep writes them
   float* d_vel = reinterpret_cast<float*>(ctx->s_acc4);    // staging; free until the force sweep writes it
   int* d_flag = &ctx->d_scalars->overflow;
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pos, pos_xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
   if (mass)
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->stage_f, mass, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
   SPH_CUDA_CHECK(ctx, cudaEventRecord(ctx->upload_ev[0], st));
   // the velocity copy follows the position copy on the link instead of sharing it
   SPH_CUDA_CHECK(ctx, cudaStreamWaitEvent(su, ctx->upload_ev[0], 0));
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_vel, vel_xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, su));
   k_pack_vel<<<blocks_for(n), kThreads, 0, su>>>(n, d_vel, ctx->vel4);
   SPH_CUDA_CHECK(ctx, cudaEventRecord(ctx->upload_ev[1], su));
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
   k_pack_pos<<<blocks_for(n), kThreads, 0, st>>>(n, d_pos, mass ? ctx->stage_f : nullptr, ctx->pos4, d_flag);
   ctx->launches += 2;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   int differs = 0;
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(&differs, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
   ctx->uniform_mass = differs == 0;
   ctx->n_local = ctx->n_owned = n;
   sph_graph_invalidate(ctx);
   ctx->voxel_ids_valid = false;
   ctx->deferred_vel_event = ctx->upload_ev[1];
   int rc = sph_step_full(ctx);
   ctx->deferred_vel_event = nullptr;
   ctx->stepped = rc == SPHB200_OK;
   if (rc == SPHB200_OK)
      ctx->steps_done++;
   return rc;
}

int sphb200_step_host(sphb200_ctx* ctx, float* pos_xyz, float* vel_xyz, const float* mass)
{
   if (!ctx || !pos_xyz || !vel_xyz)
      return sph_fail(ctx, SPHB200_E_INVALID, "step_host: null argument");
   int rc;
   if (!ctx->comm && ctx->params.neighbor_mode == SPHB200_NEIGHBORS_FULL && !ctx->params.enable_timers &&
       ctx->params.kernel_variant != 1 && ctx->capacity > 0)
   {
      SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
      rc = step_host_overlapped(ctx, pos_xyz, vel_xyz, mass);
   }
   else
   {
      rc = sphb200_upload_state(ctx, pos_xyz, vel_xyz, mass);
      if (rc)
         return rc;
      rc = sphb200_step(ctx, 1);
   }
   if (rc)
      return rc;
   const int n = ctx->n_local;
   cudaStream_t st = ctx->stream;
   if (n > 0)
   {
      float* d_pos = reinterpret_cast<float*>(ctx->s_posA4);
      float* d_vel = reinterpret_cast<float*>(ctx->s_velB4);
      k_unpack_xyz<<<blocks_for(n), kThreads, 0, st>>>(n, ctx->pos4, d_pos);
      k_unpack_xyz<<<blocks_for(n), kThreads, 0, st>>>(n, ctx->vel4, d_vel);
      ctx->launches += 2;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(pos_xyz, d_pos, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, st));
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(vel_xyz, d_vel, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, st));
   }
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
   return SPHB200_OK;
}

int sphb200_build_neighbor_lists(sphb200_ctx* ctx)
{
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null context");
   if (ctx->params.neighbor_mode != SPHB200_NEIGHBORS_FULL)
      return ctx->lists_valid ? SPHB200_OK
                              : sph_fail(ctx, SPHB200_E_INVALID, "build_neighbor_lists: step first");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   int rc = alloc_lists(ctx);
   if (rc)
      return rc;
   return sph_full_build_lists(ctx);
}

int sphb200_build_neighbor_lists_visited(sphb200_ctx* ctx)
{
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null context");
   if (ctx->params.neighbor_mode != SPHB200_NEIGHBORS_FULL)
      return sph_fail(ctx, SPHB200_E_INVALID, "build_neighbor_lists_visited: FULL neighbour mode only");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   int rc = alloc_lists(ctx);
   if (rc)
      return rc;
   return sph_full_build_lists(ctx, true);
}

int sphb200_get_energies(sphb200_ctx* ctx, float* e_kin, float* e_pot)
{
   if (!ctx || !e_kin || !e_pot)
      return sph_fail(ctx, SPHB200_E_INVALID, "null argument");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   int rc = fetch_scalars(ctx);
   if (rc == SPHB200_OK)
      rc = sph_comm_check(ctx);
   if (rc)
      return rc;
   *e_kin = (float)ctx->h_scalars.e_kin;
   *e_pot = (float)ctx->h_scalars.e_pot;
   return SPHB200_OK;
}

int sphb200_get_neighbor_stats(sphb200_ctx* ctx, long long* total, int* max_count, int* min_count)
{
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null context");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   int rc = fetch_scalars(ctx);
   if (rc == SPHB200_OK)
      rc = sph_comm_check(ctx);
   if (rc)
      return rc;
   if (total) *total = (long long)ctx->h_scalars.nbr_total;
   if (max_count) *max_count = ctx->h_scalars.nbr_max;
   if (min_count) *min_count = ctx->h_scalars.nbr_min;
   return SPHB200_OK;
}

int sphb200_get_timings(sphb200_ctx* ctx, float ms[6])
{
   if (!ctx || !ms)
      return sph_fail(ctx, SPHB200_E_INVALID, "null argument");
   for (int i = 0; i < 6; i++)
      ms[i] = 0.0f;
   if (!ctx->params.enable_timers || !ctx->stepped)
      return SPHB200_OK;
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   for (int i = 0; i < 6; i++)
   {
      float t = 0.0f;
      if (cudaEventElapsedTime(&t, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess)
         ms[i] = t;
      else
         cudaGetLastError();
   }
   return SPHB200_OK;
}

int sphb200_get_step_report(sphb200_ctx* ctx, SphStepReport* out)
{
   if (!ctx || !out)
      return sph_fail(ctx, SPHB200_E_INVALID, "null argument");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   if (!ctx->h_report)
      SPH_CUDA_CHECK(ctx, cudaMallocHost((void**)&ctx->h_report, sizeof(StepScalars)));
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->h_report, ctx->d_scalars, sizeof(StepScalars), cudaMemcpyDeviceToHost,
                                       ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   int rc = sph_comm_check(ctx);
   if (rc)
      return rc;
   ctx->h_scalars = *ctx->h_report;
   out->e_kin = (float)ctx->h_scalars.e_kin;
   out->e_pot = (float)ctx->h_scalars.e_pot;
   out->nbr_total = (long long)ctx->h_scalars.nbr_total;
   out->nbr_max = ctx->h_scalars.nbr_max;
   out->nbr_min = ctx->h_scalars.nbr_min;
   out->step_index = ctx->steps_done;
   for (int i = 0; i < 6; i++)
      out->phase_ms[i] = 0.0f;
   if (ctx->params.enable_timers && ctx->stepped)
      for (int i = 0; i < 6; i++)
      {
         float t = 0.0f;
         if (cudaEventElapsedTime(&t, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess)
            out->phase_ms[i] = t;
         else
            cudaGetLastError();
      }
   return SPHB200_OK;
}

int sphb200_snapshot_request(sphb200_ctx* ctx, int what)
{
   if (!ctx || !(what & (SPHB200_SNAP_POSITIONS | SPHB200_SNAP_CELL_COUNTS)))
      return sph_fail(ctx, SPHB200_E_INVALID, "snapshot_request: bad argument");
   if (ctx->comm)
      return sph_fail(ctx, SPHB200_E_INVALID, "snapshot_request: single-GPU contexts only");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   int rc = snapshots_init(ctx);
   if (rc)
      return rc;
   sphb200_ctx::Snapshots* sn = ctx->snap;
   // a reader copying out of a buffer right now: skip this frame instead of stalling the steps
   std::unique_lock<std::mutex> lk(sn->lock, std::try_to_lock);
   if (!lk.owns_lock())
      return SPHB200_OK;
   for (int b = 0; b < 2; b++)
      if (sn->in_flight[b] && cudaEventQuery(sn->done[b]) == cudaSuccess)
         sn->in_flight[b] = false;
   cudaGetLastError();   // cudaErrorNotReady is not an error
   // the buffer that does not hold the newest complete snapshot
   int target = sn->newest < 0 ? 0 : 1 - sn->newest;
   if (sn->in_flight[target])
      target = 1 - target;                    // the other one finished in the meantime?
   if (sn->in_flight[target] || sn->in_flight[1 - target])
      return SPHB200_OK;                      // a copy is still on the wire (one staging buffer): drop the frame
   cudaStream_t st = ctx->stream;
   const int n = ctx->n_local;
   if ((what & SPHB200_SNAP_POSITIONS) && n > 0)
   {
      k_unpack_xyz<<<blocks_for(n), kThreads, 0, st>>>(n, ctx->pos4, sn->dev_pos);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
   }
   if (what & SPHB200_SNAP_CELL_COUNTS)
   {
      rc = sph_grid_voxel_histogram(ctx, sn->dev_cnt);
      if (rc)
         return rc;
   }
   SPH_CUDA_CHECK(ctx, cudaEventRecord(sn->staged, st));
   SPH_CUDA_CHECK(ctx, cudaStreamWaitEvent(sn->stream, sn->staged, 0));
   if ((what & SPHB200_SNAP_POSITIONS) && n > 0)
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(sn->host_pos[target], sn->dev_pos, sizeof(float) * 3 * (size_t)n,
                                          cudaMemcpyDeviceToHost, sn->stream));
   if (what & SPHB200_SNAP_CELL_COUNTS)
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(sn->host_cnt[target], sn->dev_cnt, sizeof(int) * (size_t)ctx->cells_voxel,
                                          cudaMemcpyDeviceToHost, sn->stream));
   SPH_CUDA_CHECK(ctx, cudaEventRecord(sn->done[target], sn->stream));
   // the next step may overwrite pos4 as soon as the pack kernel has run (stream order); the
   // staging buffers are reused only after this copy (in_flight above)
   sn->in_flight[target] = true;
   sn->what[target] = what;
   sn->step[target] = ctx->steps_done;
   sn->newest = target;
   return SPHB200_OK;
}

int sphb200_snapshot_read(sphb200_ctx* ctx, int wait, float* pos_xyz, size_t pos_bytes, int* cell_counts,
                          size_t count_bytes, long long* step_index)
{
   if (!ctx || !step_index)
      return sph_fail(ctx, SPHB200_E_INVALID, "snapshot_read: null argument");
   sphb200_ctx::Snapshots* sn = ctx->snap;
   if (!sn || !sn->ready)
      return SPHB200_OK;                      // nothing requested yet
   // no CUDA call below touches the context's stream: safe beside a running sphb200_step
   std::lock_guard<std::mutex> lk(sn->lock);
   int b = sn->newest;
   if (b < 0)
      return SPHB200_OK;
   if (sn->in_flight[b])
   {
      if (wait)
      {
         if (cudaEventSynchronize(sn->done[b]) != cudaSuccess)
            return sph_fail(ctx, SPHB200_E_CUDA, "snapshot_read: copy failed");
         sn->in_flight[b] = false;
      }
      else if (cudaEventQuery(sn->done[b]) == cudaSuccess)
         sn->in_flight[b] = false;
      else
      {
         cudaGetLastError();
         b = 1 - b;                           // the previous snapshot, if it is complete
         if (sn->step[b] < 0 || sn->in_flight[b])
            return SPHB200_OK;
      }
   }
   if (sn->step[b] <= *step_index)
      return SPHB200_OK;                      // the caller already has it
   const size_t n = (size_t)ctx->n_local, cells = (size_t)ctx->cells_voxel;
   if (pos_xyz && (sn->what[b] & SPHB200_SNAP_POSITIONS))
   {
      if (pos_bytes < sizeof(float) * 3 * n)
         return sph_fail(ctx, SPHB200_E_INVALID, "snapshot_read: position destination too small");
      memcpy(pos_xyz, sn->host_pos[b], sizeof(float) * 3 * n);
   }
   if (cell_counts && (sn->what[b] & SPHB200_SNAP_CELL_COUNTS))
   {
      if (count_bytes < sizeof(int) * cells)
         return sph_fail(ctx, SPHB200_E_INVALID, "snapshot_read: count destination too small");
      memcpy(cell_counts, sn->host_cnt[b], sizeof(int) * cells);
   }
   *step_index = sn->step[b];
   return SPHB200_OK;
}

int sphb200_get_launch_count(const sphb200_ctx* ctx, long long* launches)
{
   if (!ctx || !launches)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null argument");
   *launches = ctx->launches;
   return SPHB200_OK;
}

}  // extern "C"
