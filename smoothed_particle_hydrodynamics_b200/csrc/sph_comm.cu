// sph_comm.cu -- multi-GPU slab decomposition along z (net-new; the reference is
// single-process).
//
// One context = one rank = one GPU owning the voxel layers [z0, z1) of the global
// grid.  z is the slowest index of computeVoxelId (sph.cpp:1151-1154), so a slab is
// a contiguous key range and local keys are global keys minus an offset.  The local
// grid also carries one ghost voxel layer (width 2h) per neighbour: with 2h of
// ghosts every particle within h of an owned particle has its own full
// neighbourhood on this rank, so its density (needed at sph.cpp:829-830) is
// recomputed locally and ONE exchange per step suffices.
//
// Storage is slot based: every slot of the particle arrays is OWNED, GHOST or FREE.
// FREE slots get a sentinel cell key and fall behind all particles in the sort, so
// nothing is ever compacted and the live count is read from the cell table on the
// device (no host round trip per step).  Per step and per neighbour one message
// {header, migrants, ghost layer} goes out and one comes in.  The message is built by
// the force sweep itself (sph_slab_emit): the kernel that integrates a boundary-layer
// particle stores it straight into the NEIGHBOUR GPU's receive buffer through a
// peer-mapped pointer (NVLink; CUDA IPC between the ranks' processes), only the
// entries that exist cross the link, and a one-thread kernel then publishes the counts
// and a message number that the neighbour's next exchange waits for.  Receive buffers
// are double buffered by message parity, which is all the flow control a symmetric
// per-step exchange needs (see slab_unpack).  Without peer access (or with
// SPHB200_HALO=nccl) the same messages are built locally and travel as one grouped
// ncclSend/ncclRecv pair of fixed capacity.  Particles carry their global id; the
// in-cell order is re-ranked by it (sph_grid.cu) so an N-slab run reproduces the
// 1-slab run.
//
// Exchange rules for an OWNED particle whose voxel layer is now vz:
//   vz >= z1 : migrant to rank+1; kept here as a GHOST when vz == z1 (it sits in the
//              neighbour's boundary layer, which this rank needs as ghost anyway)
//   vz <  z0 : symmetric, to rank-1
//   otherwise: stays; copied into the neighbour's ghost message when vz is the
//              first / last owned layer.
// Old GHOST slots are dropped every step.  A particle that crosses more than one
// slab in a step is forwarded one rank per step.
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is bound at run time, see NcclApi
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sph_math.cuh"

struct SlabComm
{
   ncclComm_t nccl;
   bool has_nccl;
   int rank, nranks;
   int z0, z1;            // owned voxel layers
   int zlo, zhi;          // local grid [zlo, zhi) = owned + ghost layers
   int mig_cap, ghost_cap;
   size_t msg_bytes;
   size_t msg_stride;        // msg_bytes rounded up to 256
   unsigned char* send[2];   // NCCL / copy mode: outgoing messages, 0: to rank-1 (down), 1: to rank+1 (up)
   unsigned char* recv_area; // one allocation: the four receive buffers [from down / from up][parity]
   unsigned char* recv[2][2];
   unsigned* flags;          // [0] / [1]: number of the last complete message from down / up (written by the peer)
   unsigned* send_cnt;       // message being built: {migrants, ghosts} down, {migrants, ghosts} up
   uint32_t* free_list;      // FREE slot indices of this step
   unsigned* counters;       // [0] n_free, [1] error flag: 1 message capacity, 2 slot capacity, 3 peer timeout
   unsigned h_counters[2];
   bool msgs_ready;          // the last force sweep already built the outgoing messages
   unsigned exchange_no;     // exchanges started so far; message M(e) is consumed by exchange e
   unsigned build_no;        // number of the message the next emit / finalize works on
   // put mode: the neighbours' receive buffers and flag words, peer mapped
   bool put_mode;
   unsigned char* peer_msg[2][2];   // [0 down / 1 up][parity]
   unsigned* peer_flag[2];
   void* ipc_base[2][2];            // what cudaIpcOpenMemHandle returned: [dir][0 receive area, 1 flags]
};

namespace
{

// NCCL is bound with dlopen at the first use instead of at link time: a process
// that also imports torch must end up with ONE libnccl.so.2 (torch bundles a newer
// one than the system's), whichever of the two libraries is loaded first.
struct NcclApi
{
   ncclResult_t (*GetUniqueId)(ncclUniqueId*);
   ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
   ncclResult_t (*CommDestroy)(ncclComm_t);
   ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
   ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
   ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
   ncclResult_t (*GroupStart)();
   ncclResult_t (*GroupEnd)();
   const char* (*GetErrorString)(ncclResult_t);
   bool ok;
   std::string why;
};

NcclApi& nccl_api()
{
   static NcclApi api = [] {
      NcclApi a;
      memset(&a, 0, offsetof(NcclApi, ok));
      a.ok = false;
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // already in the process (torch)?
      if (!h)
         if (const char* env = getenv("SPHB200_NCCL_LIB"))
            h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
      if (!h)
         h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h)
      {
         a.why = std::string("cannot load libnccl.so.2: ") + dlerror();
         return a;
      }
      *(void**)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
      *(void**)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
      *(void**)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
      *(void**)&a.Send = dlsym(h, "ncclSend");
      *(void**)&a.Recv = dlsym(h, "ncclRecv");
      *(void**)&a.AllReduce = dlsym(h, "ncclAllReduce");
      *(void**)&a.GroupStart = dlsym(h, "ncclGroupStart");
      *(void**)&a.GroupEnd = dlsym(h, "ncclGroupEnd");
      *(void**)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
      a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.AllReduce && a.GroupStart && a.GroupEnd &&
             a.GetErrorString;
      if (!a.ok)
         a.why = "libnccl.so.2 lacks a required symbol";
      return a;
   }();
   return api;
}

constexpr int kThreads = 256;

// First exchange after an upload: classify every slot, build the outgoing messages and
// the free list.  (Later exchanges get their messages from the force sweep, which has
// every particle's new position at hand -- sph_slab_emit in sph_math.cuh -- and only
// need k_slab_freelist.)
__global__ void __launch_bounds__(kThreads)
   k_slab_pack(DevParams P, int capacity, const float4* __restrict__ pos4, const float4* __restrict__ vel4,
               const uint32_t* __restrict__ gid, unsigned char* __restrict__ state,
               uint32_t* __restrict__ free_list, unsigned* __restrict__ counters)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   const bool valid = i < capacity;
   unsigned char st = valid ? state[i] : (unsigned char)SLOT_OWNED;
   const bool owned = valid && st == SLOT_OWNED;
   float4 p = make_float4(0.f, 0.f, 0.f, 0.f), v = p;
   if (owned)
   {
      p = pos4[i];
      v = vel4[i];
   }
   const unsigned char ns = sph_slab_emit(P, owned, p, v, (uint32_t)(valid ? i : 0));
   if (owned)
      st = ns == SLOT_LEAVING_GHOST ? (unsigned char)SLOT_GHOST : ns == SLOT_LEAVING_FREE ? (unsigned char)SLOT_FREE : st;
   else if (valid)
      st = SLOT_FREE;      // FREE stays free; last step's ghosts are dropped (fresh ones arrive with the unpack)
   const bool is_free = valid && st == SLOT_FREE;
   const unsigned f = sph_block_append(&counters[0], is_free);
   if (is_free)
      free_list[f] = (uint32_t)i;
   if (valid)
      state[i] = st;
}

// Exchanges whose messages were built by the force sweep: one pass over the slot states
// drops last step's ghosts, retires the migrants and lists the free slots.
__global__ void __launch_bounds__(kThreads)
   k_slab_freelist(int capacity, unsigned char* __restrict__ state, uint32_t* __restrict__ free_list,
                   unsigned* __restrict__ counters)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   const bool valid = i < capacity;
   unsigned char st = valid ? state[i] : (unsigned char)SLOT_OWNED;
   if (st == SLOT_LEAVING_GHOST)
      st = SLOT_GHOST;
   else if (st != SLOT_OWNED)
      st = SLOT_FREE;
   const bool is_free = valid && st == SLOT_FREE;
   const unsigned f = sph_block_append(&counters[0], is_free);
   if (is_free)
      free_list[f] = (uint32_t)i;
   if (valid)
      state[i] = st;
}

// arrivals of both incoming messages into free slots
__global__ void __launch_bounds__(kThreads)
   k_slab_unpack(int mig_cap, int ghost_cap, const unsigned char* __restrict__ msg_down,
                 const unsigned char* __restrict__ msg_up, const uint32_t* __restrict__ free_list,
                 unsigned* __restrict__ counters, float4* __restrict__ pos4, float4* __restrict__ vel4,
                 uint32_t* __restrict__ gid, unsigned char* __restrict__ state)
{
   // a failed exchange (peer timeout, overflow) leaves stale or partial messages behind: consume
   // nothing -- the sticky error is raised at the next host synchronisation point (sph_comm_check)
   if (counters[1] != 0u)
      return;
   const SlabMsgHeader hd = *reinterpret_cast<const SlabMsgHeader*>(msg_down);
   const SlabMsgHeader hu = *reinterpret_cast<const SlabMsgHeader*>(msg_up);
   unsigned c0 = min(hd.n_migrants, (unsigned)mig_cap), c1 = min(hd.n_ghosts, (unsigned)ghost_cap);
   unsigned c2 = min(hu.n_migrants, (unsigned)mig_cap), c3 = min(hu.n_ghosts, (unsigned)ghost_cap);
   unsigned total = c0 + c1 + c2 + c3;
   unsigned a = blockIdx.x * blockDim.x + threadIdx.x;
   if (a >= total)
      return;
   if (a >= counters[0])
   {
      atomicMax(&counters[1], 2u);   // out of free slots: particle capacity exceeded
      return;
   }
   const SlabEntry* src;
   unsigned char st;
   if (a < c0)
   {
      src = sph_msg_entries(const_cast<unsigned char*>(msg_down)) + a;
      st = SLOT_OWNED;
   }
   else if (a < c0 + c1)
   {
      src = sph_msg_entries(const_cast<unsigned char*>(msg_down)) + mig_cap + (a - c0);
      st = SLOT_GHOST;
   }
   else if (a < c0 + c1 + c2)
   {
      src = sph_msg_entries(const_cast<unsigned char*>(msg_up)) + (a - c0 - c1);
      st = SLOT_OWNED;
   }
   else
   {
      src = sph_msg_entries(const_cast<unsigned char*>(msg_up)) + mig_cap + (a - c0 - c1 - c2);
      st = SLOT_GHOST;
   }
   SlabEntry e = *src;
   uint32_t slot = free_list[a];
   pos4[slot] = e.pos;
   vel4[slot] = make_float4(e.vel.x, e.vel.y, e.vel.z, 0.0f);
   gid[slot] = __float_as_uint(e.vel.w);
   state[slot] = st;
}

// Publishes the message the previous kernel built: entry counts into its header and -- put
// mode -- the message number into the neighbour's flag word.  The entries were stored by
// the previous kernel of this stream, so they are complete (system wide) before this runs.
__global__ void k_slab_finalize(unsigned char* msg_down, unsigned char* msg_up, const unsigned* __restrict__ send_cnt,
                                unsigned* flag_down, unsigned* flag_up, unsigned msg_no)
{
   if (threadIdx.x != 0 || blockIdx.x != 0)
      return;
   if (msg_down)
   {
      SlabMsgHeader* h = reinterpret_cast<SlabMsgHeader*>(msg_down);
      h->n_migrants = send_cnt[0];
      h->n_ghosts = send_cnt[1];
   }
   if (msg_up)
   {
      SlabMsgHeader* h = reinterpret_cast<SlabMsgHeader*>(msg_up);
      h->n_migrants = send_cnt[2];
      h->n_ghosts = send_cnt[3];
   }
   __threadfence_system();
   if (flag_down)
      *reinterpret_cast<volatile unsigned*>(flag_down) = msg_no;
   if (flag_up)
      *reinterpret_cast<volatile unsigned*>(flag_up) = msg_no;
}

// put mode: waits until both neighbours have published message `want` (bounded: a peer
// that never arrives raises error 3 instead of hanging the GPU)
__device__ __forceinline__ unsigned long long sph_globaltimer_ns()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
}

__global__ void k_slab_wait(const unsigned* flags, int has_down, int has_up, unsigned want, unsigned* counters,
                            unsigned long long timeout_ns)
{
   if (threadIdx.x != 0 || blockIdx.x != 0)
      return;
   const volatile unsigned* f = flags;
   for (int d = 0; d < 2; d++)
   {
      if (!(d ? has_up : has_down))
         continue;
      const unsigned long long t0 = sph_globaltimer_ns();
      while ((int)(f[d] - want) < 0)
      {
         if (sph_globaltimer_ns() - t0 > timeout_ns)
         {
            atomicMax(&counters[1], 3u);
            break;
         }
         __nanosleep(100);
      }
   }
   __threadfence_system();
}

__global__ void __launch_bounds__(kThreads)
   k_slab_upload(int capacity, int count, const float* __restrict__ pos_xyz, const float* __restrict__ vel_xyz,
                 const float* __restrict__ mass, const uint32_t* __restrict__ ids, float4* __restrict__ pos4,
                 float4* __restrict__ vel4, uint32_t* __restrict__ gid, unsigned char* __restrict__ state,
                 int* __restrict__ mass_not_one)
{
   int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= capacity)
      return;
   if (i < count)
   {
      if (mass && mass[i] != 1.0f)
         *mass_not_one = 1;      // benign race: every writer stores the same value
      pos4[i] = make_float4(pos_xyz[3 * (size_t)i], pos_xyz[3 * (size_t)i + 1], pos_xyz[3 * (size_t)i + 2],
                            mass ? mass[i] : 1.0f);
      vel4[i] = make_float4(vel_xyz[3 * (size_t)i], vel_xyz[3 * (size_t)i + 1], vel_xyz[3 * (size_t)i + 2], 0.0f);
      gid[i] = ids[i];
      state[i] = SLOT_OWNED;
   }
   else
   {
      pos4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      vel4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      gid[i] = 0xffffffffu;
      state[i] = SLOT_FREE;
   }
}

int blocks_for(int n) { return (n + kThreads - 1) / kThreads; }

#define SPH_NCCL_CHECK(ctx, expr)                                                                        \
   do                                                                                                    \
   {                                                                                                     \
      ncclResult_t _r = (expr);                                                                          \
      if (_r != ncclSuccess)                                                                             \
         return sph_fail(ctx, SPHB200_E_COMM, std::string(#expr) + ": " + nccl_api().GetErrorString(_r)); \
   } while (0)

int require_slab(sphb200_ctx* ctx, const char* what)
{
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null context");
   if (!ctx->comm)
      return sph_fail(ctx, SPHB200_E_INVALID, std::string(what) + ": context is not a slab (call sphb200_comm_init)");
   return SPHB200_OK;
}

}  // namespace

// drops the peer mappings of put mode (closing IPC handles of other processes)
static void slab_disconnect(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   for (int d = 0; d < 2; d++)
   {
      for (int k = 0; k < 2; k++)
      {
         if (c->ipc_base[d][k])
            cudaIpcCloseMemHandle(c->ipc_base[d][k]);
         c->ipc_base[d][k] = nullptr;
      }
      c->peer_msg[d][0] = c->peer_msg[d][1] = nullptr;
      c->peer_flag[d] = nullptr;
   }
   c->put_mode = false;
}

// (re)allocates the four message buffers for `ghost_cap` ghost entries per message.
// Every rank of a run must use the SAME capacity: a send and its matching receive
// have to agree on the byte count.
static int slab_alloc_messages(sphb200_ctx* ctx, long long ghost_cap)
{
   SlabComm* c = ctx->comm;
   if (ghost_cap < 1)
      ghost_cap = 1;
   if (ghost_cap > (1ll << 28))
      return sph_fail(ctx, SPHB200_E_INVALID, "halo capacity too large");
   c->ghost_cap = (int)ghost_cap;
   c->mig_cap = (int)(ghost_cap / 4 + 256);
   c->msg_bytes = sizeof(SlabMsgHeader) + sizeof(SlabEntry) * ((size_t)c->ghost_cap + (size_t)c->mig_cap);
   c->msg_stride = (c->msg_bytes + 255) & ~(size_t)255;
   slab_disconnect(ctx);           // peer pointers would dangle
   for (int d = 0; d < 2; d++)
   {
      if (c->send[d]) cudaFree(c->send[d]);
      c->send[d] = nullptr;
      SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&c->send[d], c->msg_bytes));
      SPH_CUDA_CHECK(ctx, cudaMemset(c->send[d], 0, sizeof(SlabMsgHeader)));
   }
   if (c->recv_area) cudaFree(c->recv_area);
   c->recv_area = nullptr;
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&c->recv_area, 4 * c->msg_stride));
   for (int d = 0; d < 2; d++)
      for (int p = 0; p < 2; p++)
      {
         c->recv[d][p] = c->recv_area + (size_t)(2 * d + p) * c->msg_stride;
         SPH_CUDA_CHECK(ctx, cudaMemset(c->recv[d][p], 0, sizeof(SlabMsgHeader)));
      }
   if (!c->flags)
   {
      SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&c->flags, sizeof(unsigned) * 4));
      SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&c->send_cnt, sizeof(unsigned) * 4));
   }
   SPH_CUDA_CHECK(ctx, cudaMemset(c->flags, 0, sizeof(unsigned) * 4));
   SPH_CUDA_CHECK(ctx, cudaMemset(c->send_cnt, 0, sizeof(unsigned) * 4));
   c->exchange_no = c->build_no = 0;
   c->msgs_ready = false;
   return SPHB200_OK;
}

void sph_comm_dev_params(const sphb200_ctx* ctx, DevParams& P)
{
   const SlabComm* c = ctx->comm;
   P.slab = 1;
   P.n = ctx->capacity;
   P.gz_global = ctx->params.grid_z;
   P.vz_offset = c->zlo;
   P.gz = c->zhi - c->zlo;
   P.fz = 2 * P.gz;
   P.ghost_lo = c->z0 - c->zlo;
   P.ghost_hi = c->zhi - c->z1;
   P.slot_state = ctx->slot_state;
   P.slot_gid = ctx->gid;
   P.d_nlive = ctx->cell_start + ctx->cells_fine;
   P.own_z0 = c->z0;
   P.own_z1 = c->z1;
   P.has_down = c->rank > 0;
   P.has_up = c->rank < c->nranks - 1;
   P.mig_cap = c->mig_cap;
   P.ghost_cap = c->ghost_cap;
   // where the message under construction goes: the neighbour's receive buffer of that
   // message's parity (put mode), or the local send buffer
   const int par = (int)(c->build_no & 1u);
   P.msg_down = P.has_down ? (c->put_mode ? c->peer_msg[0][par] : c->send[0]) : nullptr;
   P.msg_up = P.has_up ? (c->put_mode ? c->peer_msg[1][par] : c->send[1]) : nullptr;
   P.send_cnt = c->send_cnt;
   P.comm_counters = c->counters;
}

// counts into the headers of the message just built and, in put mode, its number into the
// neighbours' flag words
static int slab_finalize(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   DevParams P = sph_dev_params(ctx);
   k_slab_finalize<<<1, 32, 0, ctx->stream>>>(P.msg_down, P.msg_up, c->send_cnt,
                                               c->put_mode && P.has_down ? c->peer_flag[0] : nullptr,
                                               c->put_mode && P.has_up ? c->peer_flag[1] : nullptr, c->build_no);
   ctx->launches++;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   return SPHB200_OK;
}

// start of a slab's local step: the force sweep of this step builds message exchange_no + 1
int sph_comm_begin_step(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   c->build_no = c->exchange_no + 1;
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(c->send_cnt, 0, sizeof(unsigned) * 4, ctx->stream));
   return SPHB200_OK;
}

int sph_comm_end_step(sphb200_ctx* ctx)
{
   ctx->comm->msgs_ready = true;
   return slab_finalize(ctx);
}

void sph_comm_free(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   if (!c)
      return;
   if (c->has_nccl && c->nccl)
      nccl_api().CommDestroy(c->nccl);
   slab_disconnect(ctx);
   for (int d = 0; d < 2; d++)
      if (c->send[d]) cudaFree(c->send[d]);
   if (c->recv_area) cudaFree(c->recv_area);
   if (c->flags) cudaFree(c->flags);
   if (c->send_cnt) cudaFree(c->send_cnt);
   if (c->free_list) cudaFree(c->free_list);
   if (c->counters) cudaFree(c->counters);
   if (ctx->gid) cudaFree(ctx->gid);
   if (ctx->slot_state) cudaFree(ctx->slot_state);
   if (ctx->idx_fixed) cudaFree(ctx->idx_fixed);
   ctx->gid = nullptr;
   ctx->slot_state = nullptr;
   ctx->idx_fixed = nullptr;
   delete c;
   ctx->comm = nullptr;
}

// pack this step's outgoing messages (enqueued on the context's stream)
static int slab_pack(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   cudaStream_t st = ctx->stream;
   c->exchange_no++;
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(c->counters, 0, sizeof(unsigned), st));   // [1], the error flag, is sticky
   if (c->msgs_ready)
   {
      // message exchange_no was built (and published) by the last force sweep
      if (ctx->capacity > 0)
      {
         k_slab_freelist<<<blocks_for(ctx->capacity), kThreads, 0, st>>>(ctx->capacity, ctx->slot_state, c->free_list,
                                                                         c->counters);
         ctx->launches++;
         SPH_CUDA_CHECK(ctx, cudaGetLastError());
      }
   }
   else
   {
      c->build_no = c->exchange_no;
      SPH_CUDA_CHECK(ctx, cudaMemsetAsync(c->send_cnt, 0, sizeof(unsigned) * 4, st));
      DevParams P = sph_dev_params(ctx);
      if (ctx->capacity > 0)
      {
         k_slab_pack<<<blocks_for(ctx->capacity), kThreads, 0, st>>>(P, ctx->capacity, ctx->pos4, ctx->vel4, ctx->gid,
                                                                     ctx->slot_state, c->free_list, c->counters);
         ctx->launches++;
         SPH_CUDA_CHECK(ctx, cudaGetLastError());
      }
      int rc = slab_finalize(ctx);
      if (rc)
         return rc;
   }
   c->msgs_ready = false;
   return SPHB200_OK;
}

// Arrivals of message exchange_no into free slots.  Put mode: the messages sit in the
// receive buffers of parity exchange_no & 1 once both neighbours' flags say so.  Two
// buffers suffice: a neighbour overwrites this parity again with message e + 2, during its
// force sweep of step e + 1, which it only reaches after its exchange e + 1 has seen MY
// message e + 1 -- published after my step e, i.e. after this unpack.
static int slab_unpack(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   const int par = c->put_mode ? (int)(c->exchange_no & 1u) : 0;
   if (c->put_mode)
   {
      // wall-clock bound on the wait for a neighbour (rank skew: a peer busy with a long host-side
      // download or a compile between steps): SPHB200_PEER_TIMEOUT_S, default 120 s
      double timeout_s = 120.0;
      if (const char* env = getenv("SPHB200_PEER_TIMEOUT_S"))
         if (atof(env) > 0.0)
            timeout_s = atof(env);
      k_slab_wait<<<1, 32, 0, ctx->stream>>>(c->flags, c->rank > 0, c->rank < c->nranks - 1, c->exchange_no,
                                             c->counters, (unsigned long long)(timeout_s * 1e9));
      ctx->launches++;
   }
   int max_arrivals = 2 * (c->mig_cap + c->ghost_cap);
   k_slab_unpack<<<blocks_for(max_arrivals), kThreads, 0, ctx->stream>>>(
      c->mig_cap, c->ghost_cap, c->recv[0][par], c->recv[1][par], c->free_list, c->counters, ctx->pos4, ctx->vel4,
      ctx->gid, ctx->slot_state);
   ctx->launches++;
   SPH_CUDA_CHECK(ctx, cudaGetLastError());
   ctx->voxel_ids_valid = false;
   return SPHB200_OK;
}

// pack -> (put mode: nothing, the data is already at the neighbour | one grouped
// send/recv per neighbour over NCCL) -> unpack
int sph_comm_exchange(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   if (!c->has_nccl)
      return sph_fail(ctx, SPHB200_E_COMM,
                      "step: this slab has no NCCL communicator (virtual rank): drive it with "
                      "sphb200_slab_pack / _transfer / _unpack / _step_local");
   int rc = slab_pack(ctx);
   if (rc)
      return rc;
   cudaStream_t st = ctx->stream;
   if (c->nranks > 1 && !c->put_mode)
   {
      NcclApi& N = nccl_api();
      SPH_NCCL_CHECK(ctx, N.GroupStart());
      if (c->rank > 0)
      {
         SPH_NCCL_CHECK(ctx, N.Send(c->send[0], c->msg_bytes, ncclUint8, c->rank - 1, c->nccl, st));
         SPH_NCCL_CHECK(ctx, N.Recv(c->recv[0][0], c->msg_bytes, ncclUint8, c->rank - 1, c->nccl, st));
      }
      if (c->rank < c->nranks - 1)
      {
         SPH_NCCL_CHECK(ctx, N.Send(c->send[1], c->msg_bytes, ncclUint8, c->rank + 1, c->nccl, st));
         SPH_NCCL_CHECK(ctx, N.Recv(c->recv[1][0], c->msg_bytes, ncclUint8, c->rank + 1, c->nccl, st));
      }
      SPH_NCCL_CHECK(ctx, N.GroupEnd());
   }
   return slab_unpack(ctx);
}

// Put mode between the ranks of a real run: every rank exports its receive area and flag
// words as CUDA IPC handles, neighbours swap them over the NCCL communicator and map them.
// All ranks end up in the same mode (an allreduce decides).
static int slab_connect_ipc(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   NcclApi& N = nccl_api();
   const char* env = getenv("SPHB200_HALO");
   int ok = !(env && strcmp(env, "nccl") == 0);
   struct Blob { cudaIpcMemHandle_t area, flags; };
   Blob mine, theirs[2];
   memset(&mine, 0, sizeof(mine));
   if (ok && (cudaIpcGetMemHandle(&mine.area, c->recv_area) != cudaSuccess ||
              cudaIpcGetMemHandle(&mine.flags, c->flags) != cudaSuccess))
   {
      cudaGetLastError();
      ok = 0;
   }
   unsigned char* d_blob = nullptr;      // [0] mine, [1] from down, [2] from up, then the vote
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&d_blob, 3 * sizeof(Blob) + sizeof(int)));
   cudaStream_t st = ctx->stream;
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_blob, &mine, sizeof(Blob), cudaMemcpyHostToDevice, st));
   SPH_NCCL_CHECK(ctx, N.GroupStart());
   if (c->rank > 0)
   {
      SPH_NCCL_CHECK(ctx, N.Send(d_blob, sizeof(Blob), ncclUint8, c->rank - 1, c->nccl, st));
      SPH_NCCL_CHECK(ctx, N.Recv(d_blob + sizeof(Blob), sizeof(Blob), ncclUint8, c->rank - 1, c->nccl, st));
   }
   if (c->rank < c->nranks - 1)
   {
      SPH_NCCL_CHECK(ctx, N.Send(d_blob, sizeof(Blob), ncclUint8, c->rank + 1, c->nccl, st));
      SPH_NCCL_CHECK(ctx, N.Recv(d_blob + 2 * sizeof(Blob), sizeof(Blob), ncclUint8, c->rank + 1, c->nccl, st));
   }
   SPH_NCCL_CHECK(ctx, N.GroupEnd());
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(theirs, d_blob + sizeof(Blob), 2 * sizeof(Blob), cudaMemcpyDeviceToHost, st));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
   for (int d = 0; d < 2 && ok; d++)
   {
      if (!(d ? c->rank < c->nranks - 1 : c->rank > 0))
         continue;
      void *area = nullptr, *flags = nullptr;
      if (cudaIpcOpenMemHandle(&area, theirs[d].area, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&flags, theirs[d].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      {
         cudaGetLastError();
         if (area)
            cudaIpcCloseMemHandle(area);
         ok = 0;
         break;
      }
      c->ipc_base[d][0] = area;
      c->ipc_base[d][1] = flags;
      // my message to the neighbour below arrives there as "from up" (index 1), and vice versa
      for (int p = 0; p < 2; p++)
         c->peer_msg[d][p] = static_cast<unsigned char*>(area) + (size_t)(2 * (1 - d) + p) * c->msg_stride;
      c->peer_flag[d] = static_cast<unsigned*>(flags) + (1 - d);
   }
   int* d_vote = reinterpret_cast<int*>(d_blob + 3 * sizeof(Blob));
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_vote, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
   SPH_NCCL_CHECK(ctx, N.AllReduce(d_vote, d_vote, 1, ncclInt32, ncclMin, c->nccl, st));
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(&ok, d_vote, sizeof(int), cudaMemcpyDeviceToHost, st));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
   cudaFree(d_blob);
   if (ok)
      c->put_mode = true;
   else
      slab_disconnect(ctx);
   return SPHB200_OK;
}

extern "C" {

int sphb200_comm_unique_id(void* id128)
{
   if (!id128)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null id buffer");
   static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
   NcclApi& N = nccl_api();
   if (!N.ok)
      return sph_fail(nullptr, SPHB200_E_COMM, N.why);
   ncclUniqueId id;
   ncclResult_t r = N.GetUniqueId(&id);
   if (r != ncclSuccess)
      return sph_fail(nullptr, SPHB200_E_COMM, std::string("ncclGetUniqueId: ") + N.GetErrorString(r));
   memcpy(id128, &id, 128);
   return SPHB200_OK;
}

int sphb200_comm_init(sphb200_ctx* ctx, int rank, int nranks, const void* id128, int z0, int z1)
{
   if (!ctx)
      return sph_fail(nullptr, SPHB200_E_INVALID, "null context");
   if (ctx->comm)
      return sph_fail(ctx, SPHB200_E_INVALID, "comm_init: already a slab");
   if (ctx->params.neighbor_mode != SPHB200_NEIGHBORS_FULL)
      return sph_fail(ctx, SPHB200_E_INVALID, "comm_init: slab decomposition needs neighbor_mode FULL");
   const int gz = ctx->params.grid_z;
   if (nranks < 1 || rank < 0 || rank >= nranks || z0 < 0 || z1 > gz || z1 - z0 < 1)
      return sph_fail(ctx, SPHB200_E_INVALID, "comm_init: bad rank / layer range");
   if ((rank == 0) != (z0 == 0) || (rank == nranks - 1) != (z1 == gz))
      return sph_fail(ctx, SPHB200_E_INVALID, "comm_init: slabs must tile [0, grid_z) in rank order");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   SlabComm* c = new SlabComm();
   memset(c, 0, sizeof(*c));
   c->rank = rank;
   c->nranks = nranks;
   c->z0 = z0;
   c->z1 = z1;
   c->zlo = rank > 0 ? z0 - 1 : z0;
   c->zhi = rank < nranks - 1 ? z1 + 1 : z1;
   // one voxel layer of ghosts; 2.5x the mean layer population as headroom (agreed
   // across ranks below: the largest request wins)
   long long per_layer = ((long long)ctx->capacity + (z1 - z0) - 1) / (z1 - z0);
   long long gcap = per_layer * 5 / 2 + 1024;
   if (const char* env = getenv("SPHB200_HALO_CAPACITY"))
      gcap = atoll(env);
   ctx->comm = c;
   const size_t cap = (size_t)(ctx->capacity > 0 ? ctx->capacity : 1);
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&c->free_list, sizeof(uint32_t) * cap));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&c->counters, sizeof(unsigned) * 2));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->gid, sizeof(uint32_t) * cap));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->slot_state, cap));
   SPH_CUDA_CHECK(ctx, cudaMalloc((void**)&ctx->idx_fixed, sizeof(uint32_t) * cap));
   SPH_CUDA_CHECK(ctx, cudaMemset(ctx->slot_state, SLOT_FREE, cap));
   SPH_CUDA_CHECK(ctx, cudaMemset(ctx->gid, 0xff, sizeof(uint32_t) * cap));
   // cell tables for the LOCAL grid
   sph_grid_free(ctx);
   ctx->cells_voxel = ctx->params.grid_x * ctx->params.grid_y * (c->zhi - c->zlo);
   ctx->cells_fine = 8 * ctx->cells_voxel;
   ctx->cells_alloc = ctx->cells_fine;
   int rc = sph_grid_alloc(ctx);
   if (rc)
      return rc;
   ctx->n_local = ctx->capacity;
   ctx->n_owned = 0;
   if (id128 && nranks > 1)
   {
      NcclApi& N = nccl_api();
      if (!N.ok)
         return sph_fail(ctx, SPHB200_E_COMM, N.why);
      ncclUniqueId id;
      memcpy(&id, id128, 128);
      SPH_NCCL_CHECK(ctx, N.CommInitRank(&c->nccl, nranks, id, rank));
      c->has_nccl = true;
      // all ranks must use one message size: take the largest request
      long long* d_cap = reinterpret_cast<long long*>(c->counters);   // 8 bytes of scratch
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cap, &gcap, sizeof(gcap), cudaMemcpyHostToDevice, ctx->stream));
      SPH_NCCL_CHECK(ctx, N.AllReduce(d_cap, d_cap, 1, ncclInt64, ncclMax, c->nccl, ctx->stream));
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(&gcap, d_cap, sizeof(gcap), cudaMemcpyDeviceToHost, ctx->stream));
      SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   }
   else if (nranks == 1)
      c->has_nccl = true, c->nccl = nullptr;   // single slab: nothing to exchange
   rc = slab_alloc_messages(ctx, gcap);
   if (rc)
      return rc;
   if (c->has_nccl && c->nccl)
      return slab_connect_ipc(ctx);
   return SPHB200_OK;
}

// Virtual ranks (one process): wires two neighbouring slabs for put mode -- each one's
// force sweep then writes straight into the other's receive buffers, as between GPUs.
int sphb200_slab_connect(sphb200_ctx* lower, sphb200_ctx* upper)
{
   int rc = require_slab(lower, "slab_connect");
   if (rc)
      return rc;
   rc = require_slab(upper, "slab_connect");
   if (rc)
      return rc;
   SlabComm *a = lower->comm, *b = upper->comm;
   if (b->rank != a->rank + 1 || a->msg_bytes != b->msg_bytes)
      return sph_fail(lower, SPHB200_E_INVALID, "slab_connect: contexts are not neighbours with equal halo capacity");
   if (a->exchange_no || b->exchange_no)
      return sph_fail(lower, SPHB200_E_INVALID, "slab_connect: connect before the first exchange");
   if (lower->device != upper->device)
   {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, lower->device, upper->device);
      if (!can)
         return sph_fail(lower, SPHB200_E_COMM, "slab_connect: no peer access between the two devices");
      cudaSetDevice(lower->device);
      cudaDeviceEnablePeerAccess(upper->device, 0);
      cudaSetDevice(upper->device);
      cudaDeviceEnablePeerAccess(lower->device, 0);
      cudaGetLastError();      // already enabled is fine
   }
   for (int p = 0; p < 2; p++)
   {
      a->peer_msg[1][p] = b->recv[0][p];     // lower's "up" message arrives at upper as "from down"
      b->peer_msg[0][p] = a->recv[1][p];
   }
   a->peer_flag[1] = b->flags + 0;
   b->peer_flag[0] = a->flags + 1;
   // a slab is in put mode once every neighbour it has is wired
   a->put_mode = (a->rank == 0 || a->peer_flag[0]) && a->peer_flag[1];
   b->put_mode = b->peer_flag[0] && (b->rank == b->nranks - 1 || b->peer_flag[1]);
   return SPHB200_OK;
}

int sphb200_slab_put_mode(const sphb200_ctx* ctx)
{
   return ctx && ctx->comm && ctx->comm->put_mode ? 1 : 0;
}

// virtual ranks (no communicator) cannot negotiate: the caller gives every slab the
// same halo capacity (ghost particles per message) before the first step
int sphb200_slab_set_halo_capacity(sphb200_ctx* ctx, long long ghost_particles)
{
   int rc = require_slab(ctx, "slab_set_halo_capacity");
   if (rc)
      return rc;
   if (ctx->comm->nccl)
      return sph_fail(ctx, SPHB200_E_INVALID, "slab_set_halo_capacity: ranks with a communicator agree on the "
                                              "capacity at comm_init (SPHB200_HALO_CAPACITY); their neighbours "
                                              "hold mappings of the receive buffers");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   return slab_alloc_messages(ctx, ghost_particles);
}

int sphb200_get_local_count(const sphb200_ctx* ctx, int* owned, int* ghosts)
{
   if (!ctx)
      return SPHB200_E_INVALID;
   if (!ctx->comm)
   {
      if (owned) *owned = ctx->n_owned;
      if (ghosts) *ghosts = 0;
      return SPHB200_OK;
   }
   sphb200_ctx* c = const_cast<sphb200_ctx*>(ctx);
   cudaSetDevice(ctx->device);
   std::vector<unsigned char> st((size_t)ctx->capacity);
   if (cudaMemcpyAsync(st.data(), ctx->slot_state, st.size(), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
       cudaStreamSynchronize(ctx->stream) != cudaSuccess)
      return sph_fail(c, SPHB200_E_CUDA, "get_local_count: copy failed");
   int no = 0, ng = 0;
   for (unsigned char s : st)
   {
      no += s == SLOT_OWNED || s == SLOT_LEAVING_FREE || s == SLOT_LEAVING_GHOST;
      ng += s == SLOT_GHOST;
   }
   if (owned) *owned = no;
   if (ghosts) *ghosts = ng;
   return SPHB200_OK;
}

int sphb200_upload_slab(sphb200_ctx* ctx, int count, const float* pos_xyz, const float* vel_xyz, const float* mass,
                        const uint32_t* global_ids)
{
   int rc = require_slab(ctx, "upload_slab");
   if (rc)
      return rc;
   if (count < 0 || count > ctx->capacity || (count > 0 && (!pos_xyz || !vel_xyz || !global_ids)))
      return sph_fail(ctx, SPHB200_E_INVALID, "upload_slab: bad count or null argument");
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   cudaStream_t st = ctx->stream;
   float* d_pos = reinterpret_cast<float*>(ctx->s_posA4);
   float* d_vel = reinterpret_cast<float*>(ctx->s_velB4);
   uint32_t* d_ids = reinterpret_cast<uint32_t*>(ctx->keys);
   if (count > 0)
   {
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_pos, pos_xyz, sizeof(float) * 3 * (size_t)count, cudaMemcpyHostToDevice, st));
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_vel, vel_xyz, sizeof(float) * 3 * (size_t)count, cudaMemcpyHostToDevice, st));
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(d_ids, global_ids, sizeof(uint32_t) * (size_t)count, cudaMemcpyHostToDevice, st));
      if (mass)
         SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->stage_f, mass, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice, st));
   }
   // migrants bring their own masses: the uniform-mass fast path is only safe when
   // every rank uploads unit masses (the reference's only value, sph.cpp:88); checked on the
   // device while packing
   ctx->uniform_mass = true;
   if (ctx->capacity > 0)
   {
      int* d_flag = &ctx->d_scalars->overflow;   // free between steps
      SPH_CUDA_CHECK(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
      k_slab_upload<<<blocks_for(ctx->capacity), kThreads, 0, st>>>(ctx->capacity, count, d_pos, d_vel,
                                                                    mass ? ctx->stage_f : nullptr, d_ids, ctx->pos4,
                                                                    ctx->vel4, ctx->gid, ctx->slot_state, d_flag);
      ctx->launches++;
      SPH_CUDA_CHECK(ctx, cudaGetLastError());
      int not_one = 0;
      SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(&not_one, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
      ctx->uniform_mass = not_one == 0;
   }
   ctx->n_owned = count;
   // A message built from the old state may already be published at the neighbours (put
   // mode): retire its number so that nobody consumes it.  Uploads are collective: every
   // rank of a run uploads at the same point of its step sequence (callers put a cross-rank
   // barrier before a mid-run upload: bench.py, tests/test_gpu_multiproc.py).
   // Two numbers, not one: parity is preserved, so the next message overwrites the receive buffer
   // of the dead message and never the one a lagging neighbour may still be unpacking.
   if (ctx->comm->msgs_ready)
      ctx->comm->exchange_no += 2;
   ctx->comm->msgs_ready = false;
   SPH_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->comm->counters, 0, sizeof(unsigned) * 2, st));
   ctx->lists_valid = false;
   ctx->snapshot_valid = false;
   ctx->stream_valid = false;
   ctx->voxel_ids_valid = false;
   ctx->unsorted_valid = false;
   ctx->stepped = false;
   return SPHB200_OK;
}

// OWNED particles of this slab, compacted on the host, with their global ids.
// dst_bytes is the capacity of dst; *count receives the number of particles.
int sphb200_download_slab(sphb200_ctx* ctx, int field, void* dst, size_t dst_bytes, uint32_t* global_ids, int* count)
{
   int rc = require_slab(ctx, "download_slab");
   if (rc == SPHB200_OK)
      rc = sph_comm_check(ctx);
   if (rc)
      return rc;
   if (!dst || !count)
      return sph_fail(ctx, SPHB200_E_INVALID, "download_slab: null argument");
   size_t per = 0;
   switch (field)
   {
   case SPHB200_F_POSITION: case SPHB200_F_VELOCITY: case SPHB200_F_ACCELERATION: per = 12; break;
   case SPHB200_F_MASS: case SPHB200_F_DENSITY: case SPHB200_F_NEIGHBOR_COUNT: per = 4; break;
   default:
      return sph_fail(ctx, SPHB200_E_INVALID, "download_slab: field not available per slab");
   }
   const size_t cap = (size_t)ctx->capacity;
   std::vector<unsigned char> all(per * cap), st(cap);
   std::vector<uint32_t> ids(cap);
   const int n_save = ctx->n_local;
   ctx->n_local = ctx->capacity;
   rc = sphb200_download(ctx, field, all.data(), all.size());
   ctx->n_local = n_save;
   if (rc)
      return rc;
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(st.data(), ctx->slot_state, cap, cudaMemcpyDeviceToHost, ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(ids.data(), ctx->gid, sizeof(uint32_t) * cap, cudaMemcpyDeviceToHost,
                                       ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   size_t n = 0;
   for (size_t i = 0; i < cap; i++)
      if (st[i] == SLOT_OWNED || st[i] == SLOT_LEAVING_FREE || st[i] == SLOT_LEAVING_GHOST)
      {
         if ((n + 1) * per > dst_bytes)
            return sph_fail(ctx, SPHB200_E_INVALID, "download_slab: destination too small");
         memcpy(static_cast<unsigned char*>(dst) + n * per, all.data() + i * per, per);
         if (global_ids)
            global_ids[n] = ids[i];
         n++;
      }
   *count = (int)n;
   return SPHB200_OK;
}

// ---- explicit phases (virtual ranks on one GPU; also what sphb200_step does) ----
int sphb200_slab_pack(sphb200_ctx* ctx)
{
   int rc = require_slab(ctx, "slab_pack");
   if (rc)
      return rc;
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   return slab_pack(ctx);
}

// copies src's outgoing message for direction `dir` (0 down, 1 up) into dst's matching
// receive buffer: the NCCL send/recv pair, done as a device copy between two contexts
// of one process (virtual ranks; contexts may sit on the same GPU)
int sphb200_slab_transfer(sphb200_ctx* src, int dir, sphb200_ctx* dst)
{
   int rc = require_slab(src, "slab_transfer");
   if (rc)
      return rc;
   rc = require_slab(dst, "slab_transfer");
   if (rc)
      return rc;
   if (dir < 0 || dir > 1 || src->comm->msg_bytes != dst->comm->msg_bytes ||
       dst->comm->rank != src->comm->rank + (dir ? 1 : -1))
      return sph_fail(src, SPHB200_E_INVALID, "slab_transfer: contexts are not neighbours in that direction");
   if (src->comm->put_mode || dst->comm->put_mode)
      return SPHB200_OK;      // connected slabs: the message is already in dst's receive buffer
   // the message is complete once src's stream has drained; the copy is ordered on dst's
   // stream (after dst's last unpack, before its next one).  A plain cudaMemcpy would not
   // do: device-to-device copies return before they finish and the contexts' non-blocking
   // streams do not wait for the legacy stream.
   SPH_CUDA_CHECK(src, cudaStreamSynchronize(src->stream));
   SPH_CUDA_CHECK(dst, cudaMemcpyAsync(dst->comm->recv[1 - dir][0], src->comm->send[dir], src->comm->msg_bytes,
                                       cudaMemcpyDefault, dst->stream));
   SPH_CUDA_CHECK(dst, cudaStreamSynchronize(dst->stream));   // src may repack its send buffer right away
   return SPHB200_OK;
}

int sphb200_slab_unpack(sphb200_ctx* ctx)
{
   int rc = require_slab(ctx, "slab_unpack");
   if (rc)
      return rc;
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   return slab_unpack(ctx);
}

// the local step of a slab whose exchange was driven through the explicit phases
int sphb200_slab_step_local(sphb200_ctx* ctx)
{
   int rc = require_slab(ctx, "slab_step_local");
   if (rc)
      return rc;
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   rc = sph_step_full(ctx);
   if (rc == SPHB200_OK)
      ctx->stepped = true;
   return rc;
}

// 0 = fine; otherwise SPHB200_E_CAPACITY with the reason (message or slot capacity)
int sphb200_slab_status(sphb200_ctx* ctx)
{
   int rc = require_slab(ctx, "slab_status");
   if (rc)
      return rc;
   return sph_comm_check(ctx);
}

}  // extern "C"

// The sticky error flag of the exchange, raised at every host synchronisation point of a slab
// context (synchronize, download_slab, get_energies, get_neighbor_stats, slab_status): a step
// whose exchange failed must not look like a good one to a caller that never polls the status.
int sph_comm_check(sphb200_ctx* ctx)
{
   SlabComm* c = ctx->comm;
   if (!c)
      return SPHB200_OK;
   SPH_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
   SPH_CUDA_CHECK(ctx, cudaMemcpyAsync(c->h_counters, c->counters, sizeof(unsigned) * 2, cudaMemcpyDeviceToHost,
                                       ctx->stream));
   SPH_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
   if (c->h_counters[1] == 1)
      return sph_fail(ctx, SPHB200_E_CAPACITY, "slab exchange: halo message capacity exceeded "
                                               "(raise SPHB200_HALO_CAPACITY)");
   if (c->h_counters[1] == 2)
      return sph_fail(ctx, SPHB200_E_CAPACITY, "slab exchange: no free particle slot left (raise particle_count)");
   if (c->h_counters[1] == 3)
      return sph_fail(ctx, SPHB200_E_COMM, "slab exchange: a neighbour's halo message did not arrive (peer timeout, "
                                           "SPHB200_PEER_TIMEOUT_S); the state after that step is not valid");
   return SPHB200_OK;
}
