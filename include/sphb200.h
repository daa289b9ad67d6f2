/* sphb200.h -- C ABI of the B200-native SPH step pipeline.
 *
 * Drop-in boundary for the hot path of DanielaCourel/smoothed_particle_hydrodynamics:
 * the per-timestep pipeline behind `class SPH` (reference src/sph.h:15-215,
 * src/sph.cpp:190-304).  The reference has no FFI layer -- the C++ class IS the
 * seam -- so this header is what the C++ facade
 * (smoothed_particle_hydrodynamics_b200/host/sph.h, same public interface as the
 * reference's SPH) binds to.  Plain pointers and sizes only; no torch, no Qt.
 *
 * Every entry point returns 0 on success and a negative SPHB200_E_* code on
 * failure; sphb200_last_error() gives the message.  There is NO CPU fallback:
 * without a CUDA device sphb200_create fails with SPHB200_E_CUDA.
 *
 * Threading: one context = one CUDA device + one stream; calls on a context
 * must not overlap (the reference never re-enters step() either, sph.cpp:171).
 */
#ifndef SPHB200_H
#define SPHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB200_VERSION 1

enum
{
   SPHB200_OK = 0,
   SPHB200_E_INVALID = -1,   /* bad argument / bad state                       */
   SPHB200_E_CUDA = -2,      /* CUDA runtime error or no device                */
   SPHB200_E_CAPACITY = -3,  /* particle / neighbour capacity exceeded         */
   SPHB200_E_COMM = -4       /* NCCL error (multi-GPU slabs)                   */
};

/* neighbour policy */
enum
{
   /* bit-for-bit the reference's findNeighbors sub-sampler (sph.cpp:484-692) */
   SPHB200_NEIGHBORS_REFERENCE_SAMPLED = 0,
   /* every particle within h (27 fine cells of edge h); feeds the same
    * computeDensity / computeAcceleration / integrate formulas */
   SPHB200_NEIGHBORS_FULL = 1
};

/* Every literal of SPH::SPH() (sph.cpp:36-118) plus the
 * switches the reference lacks.  sphb200_default_params() fills the reference
 * values.  Fields marked [rt] may be changed between steps with
 * sphb200_set_params (the reference's setters, sph.cpp:1225-1289); the others
 * are fixed at sphb200_create. */
typedef struct SphParams
{
   int particle_count;        /* mParticleCount (M*1024, sph.cpp:59)           */
   int grid_x, grid_y, grid_z;/* voxels per axis, edge 2h (sph.cpp:60-62)      */
   int examine_count;         /* mExamineCount E (sph.cpp:98); list capacity   */
   int neighbor_mode;         /* SPHB200_NEIGHBORS_*                            */
   int use_uniform_gravity;   /* [rt] 0 = reference (mGravity is inert, F7)    */
   int use_wall_collision;    /* [rt] 0 = reference (dead code, F6)            */
   float h;                   /* sph.cpp:47                                     */
   float simulation_scale;    /* sph.cpp:48                                     */
   float time_step;           /* [rt] sph.cpp:70                               */
   float rho0;                /* [rt] sph.cpp:74                               */
   float stiffness;           /* [rt] sph.cpp:75                               */
   float viscosity;           /* [rt] sph.cpp:77                               */
   float damping;             /* [rt] sph.cpp:78                               */
   float cfl_limit;           /* [rt] sph.cpp:89                               */
   float grav_constant;       /* [rt] sph.cpp:80                               */
   float central_mass;        /* [rt] sph.cpp:81                               */
   float central_pos[3];      /* [rt] sph.cpp:83-85; all <0 => box centre      */
   float softening;           /* [rt] sph.cpp:86;  <0 => h*scale               */
   float gravity[3];          /* [rt] sph.cpp:76                               */
   /* engine knobs (no reference counterpart) */
   int kernel_variant;        /* FULL mode: 0 = auto (tiled density sweep + flat
                                 force sweep), 1 = untiled, 3 = force sweep
                                 tiled in shared memory too (A/B)             */
   int enable_timers;         /* [rt] CUDA-event phase timers (updateElapsed)  */
   int reserved[6];
} SphParams;

/* constants the constructor derives (sph.cpp:51-57, 64-67, 90, 93-95) */
typedef struct SphDerived
{
   float h2, h_times2, h_times2_inv, h_scaled, h_scaled2, h_scaled6, h_scaled9;
   float kernel1, kernel2, kernel3;
   float cell_size, max_x, max_y, max_z;
   float cfl_limit2;
   float central_pos[3], softening;
   int grid_cell_count;
   int total_steps;           /* round(1.0 / time_step), sph.cpp:69-71         */
} SphDerived;

/* fields of sphb200_download(); layouts are the reference's host layouts */
enum
{
   SPHB200_F_POSITION = 0,        /* float[3N]  Particle::mPosition (particle.h:15)   */
   SPHB200_F_VELOCITY = 1,        /* float[3N]  Particle::mVelocity                   */
   SPHB200_F_MASS = 2,            /* float[N]   Particle::mMass                       */
   SPHB200_F_DENSITY = 3,         /* float[N]   Particle::mDensity (last step)        */
   SPHB200_F_ACCELERATION = 4,    /* float[3N]  Particle::mAcceleration (last step)   */
   SPHB200_F_NEIGHBOR_COUNT = 5,  /* int[N]     Particle::mNeighborCount              */
   SPHB200_F_VOXEL_ID = 6,        /* int[N]     SPH::mVoxelIds  (sph.h:143)           */
   SPHB200_F_VOXEL_COORD = 7,     /* int[3N]    SPH::mVoxelCoords (sph.h:144)         */
   SPHB200_F_GRID_START = 8,      /* int[cells+1] CSR offsets of SPH::mGrid lists     */
   SPHB200_F_GRID_MEMBERS = 9,    /* uint32[N]  mGrid[c] contents, in push_back order */
   SPHB200_F_CELL_COUNT = 10,     /* int[cells] mGrid[c].count() (visualization.cpp:193) */
   SPHB200_F_NEIGHBOR_INDEX = 11, /* uint32[N*E] SPH::mNeighbors (sph.h:173)          */
   SPHB200_F_NEIGHBOR_DISTANCE = 12, /* float[N*E] SPH::mNeighborDistancesScaled      */
   SPHB200_F_FINE_KEY = 13        /* int[N]     FULL mode fine-cell key               */
};

typedef struct sphb200_ctx sphb200_ctx;

/* ---- lifetime: SPH::SPH() / ~SPH() (sph.cpp:36-125) ------------------------ */
int sphb200_default_params(SphParams* p);
int sphb200_derive(const SphParams* p, SphDerived* out);
/* device < 0: current device.  Allocates all HBM buffers; state is zero. */
int sphb200_create(const SphParams* p, int device, sphb200_ctx** out);
int sphb200_destroy(sphb200_ctx* ctx);
const char* sphb200_last_error(const sphb200_ctx* ctx);   /* ctx may be NULL */

/* ---- configuration: the setters of sph.cpp:1225-1289 ----------------------- */
int sphb200_get_params(const sphb200_ctx* ctx, SphParams* out);
int sphb200_set_params(sphb200_ctx* ctx, const SphParams* p);   /* [rt] fields only */
int sphb200_get_derived(const sphb200_ctx* ctx, SphDerived* out);
/* run on a caller-owned CUDA stream (cudaStream_t as void*); NULL = own stream */
int sphb200_set_stream(sphb200_ctx* ctx, void* cuda_stream);

/* ---- scenes (host generators; initParticlePolitionsSphere, sph.cpp:361-425) */
/* the constructor's seeded rotating sphere, bit-identical (glibc rand()) */
int sphb200_scene_sphere(const SphParams* p, float* pos_xyz, float* vel_xyz);
/* jittered cubic lattice (SURVEY 8(d) scene rule), ids first_id..first_id+count-1 */
int sphb200_scene_lattice(int nx, int ny, int nz, float spacing, const float origin[3],
                          uint32_t seed, long long first_id, long long count, float* pos_xyz);

/* Named throughput scenes (BASELINE.json configs 2-5; SURVEY 8(d) scene rule): a block of
 * fluid on the jittered lattice of spacing d = h (4 pi / (3 nu))^(1/3), h = 0.1, at rest,
 * under uniform gravity (0,-9.8,0) in a box with reflecting walls -- the two switches the
 * reference's GUI rows `gravity` / `damping` were meant to drive (sphconfig.cpp:76-95) but its
 * physics never reads. */
enum
{
   SPHB200_SCENE_DAMBREAK_16K = 0,   /*  32 x  16 x  32 sites, box  20 x  8 x   8 voxels (parity tests)  */
   SPHB200_SCENE_DAMBREAK_128K = 1,  /*  64 x  32 x  64 sites, box  40 x 16 x  16 voxels                 */
   SPHB200_SCENE_DAMBREAK_1M = 2,    /* 128 x  64 x 128 sites, box  80 x 32 x  32 voxels (config 2)      */
   SPHB200_SCENE_DAMBREAK_16M = 3,   /* 256 x 128 x 512 sites, box 160 x 64 x 128 voxels (config 3)      */
   SPHB200_SCENE_BOXDROP_16M = 4     /* the same block lifted off the floor and centred (config 4, per GPU) */
};

typedef struct SphSceneLattice
{
   int nx, ny, nz;            /* lattice sites; particle id = (z*ny + y)*nx + x          */
   float spacing;             /* d                                                       */
   float origin[3];           /* min corner of the block                                 */
   uint32_t seed;             /* jitter hash seed (42)                                   */
} SphSceneLattice;

/* Fills `p` (from sphb200_default_params values: particle count, voxel grid -- grown when a
 * sparse lattice, nu < 40, is taller than the nominal box --, FULL neighbour mode, uniform
 * gravity + walls on, central mass off, rest density 1/d^3 of the lattice, examine_count
 * sized for nu) and the lattice descriptor of the scene at `nu` mean neighbours (40 = the
 * "default smoothing radius" of configs 2-4; 30 / 60 / 120 = config 5). */
int sphb200_scene_config(int scene, float nu, SphParams* p, SphSceneLattice* lattice);
/* positions (and zero velocities, when vel_xyz != NULL) of particles first_id ..
 * first_id + count - 1 of a configured scene */
int sphb200_scene_generate(const SphSceneLattice* lattice, long long first_id, long long count,
                           float* pos_xyz, float* vel_xyz);

/* ---- state: host <-> HBM (Particle arrays, particle.h:13-18) ----------------- */
/* xyz-interleaved host arrays of particle_count entries; mass may be NULL (=1) */
int sphb200_upload_state(sphb200_ctx* ctx, const float* pos_xyz, const float* vel_xyz, const float* mass);
int sphb200_download(sphb200_ctx* ctx, int field, void* dst, size_t dst_bytes);

/* ---- the hot path: SPH::step() (sph.cpp:190-304) ---------------------------- */
/* n_steps device-resident steps; asynchronous on the context's stream */
int sphb200_step(sphb200_ctx* ctx, int n_steps);
int sphb200_synchronize(sphb200_ctx* ctx);
/* one reference-style step on HOST buffers: upload, step, download pos+vel
 * (what SPH::step() does to Particle::mPosition / mVelocity in place) */
int sphb200_step_host(sphb200_ctx* ctx, float* pos_xyz, float* vel_xyz, const float* mass);
/* FULL mode only: materialise mNeighbors / mNeighborDistancesScaled for the
 * LAST step's pre-step positions (the hot kernels never store lists) */
int sphb200_build_neighbor_lists(sphb200_ctx* ctx);
/* The same lists rebuilt from what the HOT PATH itself recorded: the hit-mask stream the
 * density sweep of the last step wrote and the force sweep walked (exact test of
 * sph.cpp:633-653 applied to the survivors, in the sweep's visiting order).  A parity
 * instrument: equal to the independent scan above <=> the stream-driven force sweep saw
 * exactly the reference's neighbour sets.  FULL mode, after a step. */
int sphb200_build_neighbor_lists_visited(sphb200_ctx* ctx);

/* ---- per-step scalars ------------------------------------------------------- */
/* mKineticEnergyTotal / mPotentialEnergyTotal of the last step (sph.cpp:1004-1007) */
int sphb200_get_energies(sphb200_ctx* ctx, float* e_kin, float* e_pot);
/* sum / max / min of mNeighborCount of the last step (sph.cpp:226-232) */
int sphb200_get_neighbor_stats(sphb200_ctx* ctx, long long* total, int* max_count, int* min_count);
/* ms of the last step: voxelize, findNeighbors, density, pressure, acceleration,
 * integrate -- the six slots of SPH::updateElapsed (sph.h:73-81) */
int sphb200_get_timings(sphb200_ctx* ctx, float ms[6]);
/* everything above in ONE call and one synchronisation (what SPH::step() of the facade needs
 * after every step: sph.cpp:226-232, 292-299, 1001-1008) */
typedef struct SphStepReport
{
   float e_kin, e_pot;
   long long nbr_total;
   int nbr_max, nbr_min;
   float phase_ms[6];          /* zeros unless enable_timers                              */
   long long step_index;       /* steps this context has run since creation               */
} SphStepReport;
int sphb200_get_step_report(sphb200_ctx* ctx, SphStepReport* out);

/* ---- viewer snapshots: the readback contract of the GL view -----------------------------
 * Visualization::paintGL runs on a 16 ms timer (visualization.cpp:24-33) and reads
 * getParticles()->mPosition[3i+k] for every particle (137-163) and getGrid()[c].count() for
 * every voxel (166-213), from the GUI thread, while the worker thread steps.  With the state
 * in HBM that becomes an asynchronous snapshot: sphb200_snapshot_request (stepping thread)
 * packs the positions (and per-voxel counts) behind the steps already submitted and starts a
 * device -> pinned-host copy on a side stream -- the step stream does not wait for PCIe --;
 * sphb200_snapshot_read (ANY thread, also while the stepping thread is inside sphb200_step)
 * copies the newest COMPLETED snapshot into the caller's arrays.  Two pinned buffers; a
 * request that finds its buffer still in flight or being read is dropped (a skipped frame),
 * never waited for.  Single-GPU contexts only. */
enum
{
   SPHB200_SNAP_POSITIONS = 1,    /* float[3N], Particle::mPosition layout                 */
   SPHB200_SNAP_CELL_COUNTS = 2   /* int[cells], mGrid[c].count() of the current positions */
};
int sphb200_snapshot_request(sphb200_ctx* ctx, int what);
/* *step_index: in = the snapshot the caller already has (-1: none), out = the one copied.
 * Returns SPHB200_OK with *step_index unchanged when there is nothing newer.  wait != 0
 * blocks until the most recently requested snapshot is complete.  Either destination may be
 * NULL. */
int sphb200_snapshot_read(sphb200_ctx* ctx, int wait, float* pos_xyz, size_t pos_bytes, int* cell_counts,
                          size_t count_bytes, long long* step_index);

/* kernels launched by this context since creation (bench.py's gpu_launches) */
int sphb200_get_launch_count(const sphb200_ctx* ctx, long long* launches);

/* ---- multi-GPU slabs along z (net-new) -------------------------
 * One context per GPU owns the voxel layers [z0, z1) of the GLOBAL grid given in
 * SphParams (grid_z = whole box); particle_count is the slab's slot CAPACITY
 * (owned + ghost particles + headroom for migration).  FULL neighbour mode only.
 * In a slab context sphb200_step() first exchanges migrants and the ghost layer
 * with the z-neighbours (peer puts over NVLink by the force sweep of the previous step,
 * or one grouped ncclSend/ncclRecv pair each) and then runs the local step. */
/* 128-byte NCCL unique id (rank 0 makes it, the caller broadcasts it through its
 * own channel, e.g. torch.distributed) */
int sphb200_comm_unique_id(void* id128);
/* id128 == NULL makes a "virtual rank" without a communicator: several slabs in
 * one process (even on one GPU) exchanged through sphb200_slab_transfer */
int sphb200_comm_init(sphb200_ctx* ctx, int rank, int nranks, const void* id128, int z0, int z1);
int sphb200_get_local_count(const sphb200_ctx* ctx, int* owned, int* ghosts);
/* `count` owned particles with their global ids; the remaining slots are free.  Collective:
 * every rank of a run uploads at the same point of its step sequence (a halo message built
 * from the old state is retired by number on all ranks alike). */
int sphb200_upload_slab(sphb200_ctx* ctx, int count, const float* pos_xyz, const float* vel_xyz,
                        const float* mass, const uint32_t* global_ids);
/* owned particles only, compacted, with their global ids (any order).  Fields:
 * POSITION, VELOCITY, MASS, DENSITY, ACCELERATION, NEIGHBOR_COUNT */
int sphb200_download_slab(sphb200_ctx* ctx, int field, void* dst, size_t dst_bytes, uint32_t* global_ids,
                          int* count);
/* the phases of the exchange, for virtual ranks: pack on every slab, transfer each
 * message to the neighbour (dir 0 = to rank-1, 1 = to rank+1), unpack, step_local */
int sphb200_slab_pack(sphb200_ctx* ctx);
int sphb200_slab_transfer(sphb200_ctx* src, int dir, sphb200_ctx* dst);
int sphb200_slab_unpack(sphb200_ctx* ctx);
int sphb200_slab_step_local(sphb200_ctx* ctx);
/* ghost particles per halo message.  NCCL ranks agree on the largest request at
 * comm_init; virtual ranks must be given one common value by the caller. */
int sphb200_slab_set_halo_capacity(sphb200_ctx* ctx, long long ghost_particles);
/* Put mode for two neighbouring virtual ranks (upper.rank == lower.rank + 1), before their
 * first exchange: each slab's force sweep then stores its halo particles and migrants
 * straight into the other's receive buffers (peer memory), which is what real ranks do
 * over NVLink after sphb200_comm_init (CUDA IPC; SPHB200_HALO=nccl keeps the grouped
 * ncclSend/ncclRecv instead).  sphb200_slab_transfer becomes a no-op for connected slabs. */
int sphb200_slab_connect(sphb200_ctx* lower, sphb200_ctx* upper);
/* In put mode a slab's receive buffers are written by its neighbours: destroy the contexts
 * of a run together (after a barrier of the caller's), not while a neighbour still steps. */
/* 1 when the slab exchanges by peer puts, 0 when by NCCL / sphb200_slab_transfer */
int sphb200_slab_put_mode(const sphb200_ctx* ctx);
/* SPHB200_E_CAPACITY when a halo message or the slot capacity overflowed */
int sphb200_slab_status(sphb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H */
