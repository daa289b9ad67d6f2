// sph_internal.h -- context and device-parameter block shared by the .cu files.
#ifndef SPHB200_INTERNAL_H
#define SPHB200_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "sphb200.h"

// Everything a kernel needs, passed by value (fits the 4 KB parameter space).
// Derived constants are computed on the host with the reference constructor's
// own expressions (sph.cpp:47-95) so that they carry the same rounding.
struct DevParams
{
   int n;                 // particles on this device (owned + ghosts in slab mode)
   int n_owned;           // particles that are integrated / written back
   int gx, gy, gz;        // voxel grid (edge 2h)
   int fx, fy, fz;        // fine grid (edge h) = 2 * voxel grid
   int examine;           // E
   int use_gravity, use_walls;
   float h, h2, h_times2, h_times2_inv;
   float hs, hs2;         // h * scale, (h*scale)^2
   float k1, k2, k3;      // poly6, spiky-gradient, viscosity-laplacian coefficients
   float scale;           // mSimulationScale
   float dt, pos_dt;      // mTimeStep, mTimeStep * (1/scale)  (sph.cpp:956)
   float rho0, stiffness, viscosity, damping;
   float cfl, cfl2;
   float neg_gm;          // (-G) * M   (sph.cpp:913, 987)
   float gm;              // G * M      (sph.cpp:1007)
   float cx, cy, cz;      // central mass position
   float softening;
   float gvx, gvy, gvz;   // uniform gravity
   float max_x, max_y, max_z;
   int defer_velocity;    // density sweep leaves the velocity part of the force records to k_gather_vel
   // slab mode (one z-slab of the global grid per GPU; all zero / null otherwise)
   int slab;              // 1: slots may be FREE, live count comes from the device cell table
   int gz_global;         // voxel layers of the whole box (binning clamps against this, like the reference)
   int vz_offset;         // first voxel layer of the local grid (global layer index)
   int ghost_lo, ghost_hi;// ghost voxel layers present below / above the owned range (0 or 1)
   unsigned char* slot_state;         // per slot: SLOT_*
   const uint32_t* slot_gid;          // per slot: global particle id
   const uint32_t* d_nlive;           // live (non-FREE) particles = cell_start[cells]
   // halo messages of the NEXT exchange, appended to by the force sweep (sph_comm.cu)
   int own_z0, own_z1;                // owned voxel layers (global layer indices)
   int has_down, has_up;              // neighbour ranks present
   int mig_cap, ghost_cap;            // message capacities (entries)
   unsigned char* msg_down;           // to rank - 1: local send buffer, or the neighbour's receive buffer (peer mapped)
   unsigned char* msg_up;             // to rank + 1
   unsigned* send_cnt;                // entries so far: {migrants, ghosts} down, {migrants, ghosts} up
   unsigned* comm_counters;           // [0] free slots of this exchange, [1] error flag
};

// SLOT_LEAVING_*: an OWNED particle the force sweep has already put into a migrant
// message; it still counts as owned until the next exchange turns the slot into a
// FREE slot / a GHOST (it sits in the neighbour's boundary layer, needed here as ghost).
enum { SLOT_FREE = 0, SLOT_OWNED = 1, SLOT_GHOST = 2, SLOT_LEAVING_FREE = 3, SLOT_LEAVING_GHOST = 4 };

struct SlabMsgHeader
{
   unsigned n_migrants, n_ghosts, pad0, pad1;
};

struct SlabEntry      // 32 bytes per particle on the wire
{
   float4 pos;        // x, y, z, mass
   float4 vel;        // vx, vy, vz, global id (bits)
};

// number of particles a kernel has to process: by value on a single GPU, read from
// the cell table (no host round trip) in slab mode
__device__ __forceinline__ int sph_live_count(const DevParams& P)
{
   return P.slab ? (int)*P.d_nlive : P.n;
}

struct StepScalars     // per-step reductions, one device struct
{
   double e_kin, e_pot;
   unsigned long long nbr_total;
   int nbr_max, nbr_min;
   int overflow;        // FULL list builder: count above capacity seen
   unsigned finish_ticket;   // k_finish_scalars: blocks that have delivered their partial sum (self-resetting)
};

struct SlabComm;       // sph_comm.cu

struct sphb200_ctx
{
   SphParams params;
   SphDerived derived;
   int device;
   cudaStream_t stream;
   bool own_stream;
   std::string error;
   long long launches;

   int capacity;          // allocated particle slots
   int n_local;           // particles currently on this device (== capacity single GPU)
   int n_owned;
   int cells_voxel, cells_fine, cells_alloc;

   // persistent state, original (upload) order: xyz+mass, vel+pad
   float4* pos4;
   float4* vel4;
   uint32_t* gid;         // slab mode: global id per slot, else NULL
   unsigned char* slot_state;   // slab mode: SLOT_* per slot, else NULL
   uint32_t* idx_fixed;   // slab mode: idx_sorted re-ordered by global id inside each cell

   // per-step scratch
   uint32_t *keys, *keys_sorted, *idx_iota, *idx_sorted;
   const uint32_t* idx_order;   // particle order of the last binning: idx_sorted, or idx_fixed in slab mode
   uint32_t* cell_slot;   // per particle: its slot inside its cell (return value of the histogram atomic)
   unsigned long long* tmp_pair;  // counting-sort scatter output per position: (x order key << 32 | id), the value the
                                  // members of a cell are ranked by (x key 0 in sampled mode: ascending id = push_back order)
   uint32_t* tmp_idx;             // slab mode: the slot index per scattered position (id = global id there)
   uint32_t* xkey;                // FULL mode: order key of x per particle (in-cell order by x)
   bool use_radix_sort;   // A/B switch (env SPHB200_RADIX_SORT=1): CUB radix sort instead of the counting sort
   uint32_t* cell_count;  // histogram, cells_alloc + 1
   uint32_t* cell_start;  // exclusive scan, cells_alloc + 1
   float4* s_pos4;        // sorted snapshot (x,y,z,m)
   float4* s_posA4;       // sorted (x,y,z,fA)   force-pass record
   float4* s_velB4;       // sorted (vx,vy,vz,fB)
   float* s_rho;          // sorted density
   float4* s_acc4;        // sorted acceleration (x,y,z,unused)
   int* s_count;          // sorted neighbour count
   uint2* hit_rec;        // FULL mode hit-mask stream {mask, smem byte offset}, see sph_full.cu
   unsigned* hit_info;    // per sorted particle: records | hits << 8, or 0xff = scan
   uint32_t* tile_list;   // density sweep: the non-empty 8x8x4 tiles of this step (x | y << 10 | z << 20, in tiles)
   int* tile_ctl;         // [0] tiles in the list, [1] next tile to hand out
   int sm_count;
   cudaTextureObject_t tex_posA, tex_velB;   // linear float4 views of s_posA4 / s_velB4 (force-sweep A/B: TEX path)

   // SAMPLED mode / on-demand lists, original order
   uint32_t* nbr_idx;     // N * E
   float* nbr_dist;       // N * E
   int* nbr_count;        // N
   float* rho;            // N  (original order)
   float4* acc4;          // N  (original order)
   int* voxel_id;         // N  (original order)
   bool uniform_mass;     // every particle has the same mass (checked at upload): density fast path
   bool lists_valid;
   bool voxel_ids_valid;  // voxel_id[] matches the last binning / current upload
   bool snapshot_valid;   // sorted snapshot + cell table of the last FULL step present
   bool stream_valid;     // hit-mask stream of the last FULL step present (tiled density sweep)
   uint32_t *vg_count, *vg_start, *vg_members, *vg_keys;   // lazily allocated voxel-grid views

   void* cub_temp;
   size_t cub_temp_bytes;
   int key_bits;

   StepScalars* d_scalars;
   StepScalars h_scalars;
   double* d_block_partials;   // per-block (ekin, epot) partials
   int partial_blocks;

   float* stage_f;             // N floats: mass staging for upload / download
   bool unsorted_valid;        // rho / acc4 / nbr_count hold the last FULL step, particle order

   cudaStream_t upload_stream;       // sphb200_step_host: velocity upload beside the first half of the step
   cudaEvent_t upload_ev[2];         // [0] positions on the device, [1] velocities packed
   cudaEvent_t deferred_vel_event;   // non-null while such a step is being enqueued
   cudaGraphExec_t graph_exec; // one captured step (sph_capi.cu: step_graph), or null
   long long graph_launches;   // kernels per replay
   bool use_graph;             // env SPHB200_NO_GRAPH=1 turns the replay off (A/B)

   cudaEvent_t ev[8];
   float phase_ms[6];
   bool stepped;

   SlabComm* comm;

   long long steps_done;       // steps run since creation (snapshot / report stamps)
   StepScalars* h_report;      // pinned copy of d_scalars (sphb200_get_step_report)

   // viewer snapshots (sphb200_snapshot_request / _read)
   struct Snapshots
   {
      std::mutex lock;            // buffer roles below; read() holds it while copying out
      cudaStream_t stream;        // side stream of the D2H copies
      cudaEvent_t staged;         // staging buffers written (step stream)
      cudaEvent_t done[2];        // copy into pinned buffer b complete (side stream)
      float* dev_pos;             // float[3 * capacity] staging
      uint32_t* dev_cnt;          // uint32[cells_voxel] staging
      float* host_pos[2];         // pinned
      int* host_cnt[2];           // pinned
      int what[2];
      long long step[2];          // step index of the snapshot in buffer b (-1: empty)
      bool in_flight[2];
      int newest;                 // buffer of the most recent request (-1: none)
      bool ready;
   } * snap;
};

#define SPH_CUDA_CHECK(ctx, expr)                                                        \
   do                                                                                    \
   {                                                                                     \
      cudaError_t _e = (expr);                                                           \
      if (_e != cudaSuccess)                                                             \
         return sph_fail(ctx, SPHB200_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
   } while (0)

int sph_fail(sphb200_ctx* ctx, int code, const std::string& msg);
void sph_graph_invalidate(sphb200_ctx* ctx);
DevParams sph_dev_params(const sphb200_ctx* ctx);

// sph_grid.cu
int sph_bin_and_sort(sphb200_ctx* ctx, bool fine);
int sph_download_grid(sphb200_ctx* ctx, int field, void* dst, size_t bytes);
int sph_refresh_voxel_ids(sphb200_ctx* ctx);
int sph_grid_voxel_histogram(sphb200_ctx* ctx, uint32_t* d_counts);   // per-voxel particle counts of the current positions
int sph_grid_setup(sphb200_ctx* ctx);
// sph_sampled.cu
int sph_step_sampled(sphb200_ctx* ctx);
// sph_full.cu
int sph_step_full(sphb200_ctx* ctx);
int sph_full_build_lists(sphb200_ctx* ctx, bool from_stream = false);
int sph_full_configure(sphb200_ctx* ctx);
int sph_full_tile_count(const sphb200_ctx* ctx);
// sph_reduce.cu (in sph_grid.cu)
int sph_reset_scalars(sphb200_ctx* ctx);
int sph_finish_scalars(sphb200_ctx* ctx, int blocks);
// sph_comm.cu
int sph_comm_exchange(sphb200_ctx* ctx);
int sph_comm_begin_step(sphb200_ctx* ctx);
int sph_comm_end_step(sphb200_ctx* ctx);
void sph_comm_free(sphb200_ctx* ctx);
int sph_comm_check(sphb200_ctx* ctx);   // sticky exchange error -> status code (no-op without a slab)
void sph_comm_dev_params(const sphb200_ctx* ctx, DevParams& P);
int sph_grid_alloc(sphb200_ctx* ctx);
void sph_grid_free(sphb200_ctx* ctx);

#endif
