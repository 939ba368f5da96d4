// multi.cu -- the multi-GPU entry points of the C-ABI.  The hot path shards into independent units
// (frames of a batch, row bands of one very large image) with NO exchange step; what a host needs across GPUs is
//   * zos_multi_launch / zos_multi_sync: one host thread starts the programs of several contexts (one per device)
//     without waiting for any of them, then waits for all;
//   * an optional gather of the outputs:
//       - zos_gather_peer: contexts of ONE process, copy engines over NVLink (cudaMemcpyPeerAsync), ordered after the
//         producing kernels on each source stream and before anything later on the destination stream;
//       - zos_gather_nccl: one process per GPU.  NCCL is resolved at run time (dlopen "libnccl.so.2": the library a
//         torch.distributed host already has in the process, or the system one) so libzosimos_cuda.so has no link-time
//         dependency on it; equal shards to every rank use ncclAllGather (NVLS capable), anything else a group of
//         ncclSend / ncclRecv.  Everything is enqueued on the context's stream: no host synchronisation.
// The reference has nothing comparable: its Pool may hold several devices but always picks the first
// (lib/zosimos/src/pool.rs:227-230; "If multi-device then this should become a set", run.rs:416-420).
#include <dlfcn.h>
#include <string.h>

#include "zos_internal.h"

namespace {

// ---- the few NCCL entry points used, by their public C signatures (nccl.h 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;  // 0 = ncclSuccess
enum { ncclUint8 = 1 };    // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string why;
};

Nccl* nccl() {
  static Nccl n;
  static bool tried = false;
  if (tried) return &n;
  tried = true;
  const char* env = getenv("ZOS_NCCL_LIBRARY");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.lib) break;
  }
  if (!n.lib) { n.why = "libnccl.so.2 not found (set ZOS_NCCL_LIBRARY)"; return &n; }
#define ZOS_SYM(field, name)                                              \
  *(void**)(&n.field) = dlsym(n.lib, name);                               \
  if (!n.field) { n.why = std::string("missing symbol ") + name; dlclose(n.lib); n.lib = nullptr; return &n; }
  ZOS_SYM(GetUniqueId, "ncclGetUniqueId") ZOS_SYM(CommInitRank, "ncclCommInitRank") ZOS_SYM(CommDestroy, "ncclCommDestroy")
  ZOS_SYM(AllGather, "ncclAllGather") ZOS_SYM(Send, "ncclSend") ZOS_SYM(Recv, "ncclRecv") ZOS_SYM(GroupStart, "ncclGroupStart")
  ZOS_SYM(GroupEnd, "ncclGroupEnd") ZOS_SYM(GetErrorString, "ncclGetErrorString") ZOS_SYM(GetVersion, "ncclGetVersion")
#undef ZOS_SYM
  return &n;
}

}  // namespace

struct zos_comm {
  zos_ctx* ctx;
  ncclComm_t comm;
  uint32_t rank, world;
};

namespace {
zos_status nccl_check(zos_ctx* ctx, ncclResult_t r, const char* what) {
  if (r == 0) return ZOS_OK;
  return zos::fail(ctx, ZOS_ERR_CUDA, "%s: NCCL error %d (%s)", what, r, nccl()->GetErrorString ? nccl()->GetErrorString(r) : "?");
}
bool in_buf(const zos_buf* b, uint64_t off, uint64_t bytes) { return b && off <= b->size && bytes <= b->size - off; }
}  // namespace

extern "C" {

// ------------------------------------------------------------------ one host thread, several contexts
zos_status zos_multi_launch(zos_program* const* progs, uint32_t n, uint32_t flags) {
  if (!progs) return ZOS_ERR_INVALID;
  for (uint32_t i = 0; i < n; i++) {
    if (!progs[i]) return ZOS_ERR_INVALID;
    zos_status st = zos_program_run(progs[i], flags);  // enqueues on that context's stream and returns
    if (st != ZOS_OK) return st;
  }
  return ZOS_OK;
}

zos_status zos_multi_sync(zos_ctx* const* ctxs, uint32_t n) {
  if (!ctxs) return ZOS_ERR_INVALID;
  zos_status first = ZOS_OK;
  for (uint32_t i = 0; i < n; i++) {
    if (!ctxs[i]) return ZOS_ERR_INVALID;
    zos_status st = zos_sync(ctxs[i]);
    if (st != ZOS_OK && first == ZOS_OK) first = st;
  }
  return first;
}

zos_status zos_gather_peer(zos_ctx* dst_ctx, zos_buf* dst, const uint64_t* dst_offsets, zos_ctx* const* src_ctxs, const zos_buf* const* srcs,
                           const uint64_t* src_offsets, const uint64_t* bytes, uint32_t n) {
  if (!dst_ctx || !dst || !dst_offsets || !src_ctxs || !srcs || !bytes) return ZOS_ERR_INVALID;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t so = src_offsets ? src_offsets[i] : 0;
    if (!src_ctxs[i] || !in_buf(srcs[i], so, bytes[i]) || !in_buf(dst, dst_offsets[i], bytes[i]))
      return zos::fail(dst_ctx, ZOS_ERR_INVALID, "gather_peer: shard %u out of bounds", i);
  }
  for (uint32_t i = 0; i < n; i++) {
    if (bytes[i] == 0) continue;
    zos_ctx* sc = src_ctxs[i];
    const uint64_t so = src_offsets ? src_offsets[i] : 0;
    // pushed by the source's stream (after the kernels that produced the shard); the destination stream waits for it
    cudaSetDevice(sc->device);
    if (sc->device != dst_ctx->device) {  // direct NVLink path (without it the copy is staged through host memory)
      int can = 0;
      cudaDeviceCanAccessPeer(&can, sc->device, dst_ctx->device);
      if (can && cudaDeviceEnablePeerAccess(dst_ctx->device, 0) != cudaSuccess) cudaGetLastError();  // already enabled
    }
    cudaError_t e = sc->device == dst_ctx->device
                        ? cudaMemcpyAsync((uint8_t*)dst->ptr + dst_offsets[i], (const uint8_t*)srcs[i]->ptr + so, bytes[i], cudaMemcpyDeviceToDevice, sc->stream)
                        : cudaMemcpyPeerAsync((uint8_t*)dst->ptr + dst_offsets[i], dst_ctx->device, (const uint8_t*)srcs[i]->ptr + so, sc->device, bytes[i], sc->stream);
    zos_status st = zos::check_cuda(dst_ctx, e, "gather_peer copy");
    if (st != ZOS_OK) return st;
    if (sc != dst_ctx) {
      cudaEvent_t ev;
      if ((st = zos::check_cuda(dst_ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "gather_peer event")) != ZOS_OK) return st;
      cudaEventRecord(ev, sc->stream);
      cudaSetDevice(dst_ctx->device);
      e = cudaStreamWaitEvent(dst_ctx->stream, ev, 0);
      cudaEventDestroy(ev);  // released once the wait has been satisfied
      if ((st = zos::check_cuda(dst_ctx, e, "gather_peer wait")) != ZOS_OK) return st;
    }
  }
  cudaSetDevice(dst_ctx->device);
  return ZOS_OK;
}

// ------------------------------------------------------------------ one process per GPU: NCCL
zos_status zos_comm_unique_id(uint8_t* id128) {
  if (!id128) return ZOS_ERR_INVALID;
  Nccl* n = nccl();
  if (!n->lib) return zos::fail(nullptr, ZOS_ERR_UNSUPPORTED, "NCCL unavailable: %s", n->why.c_str());
  ncclUniqueId id;
  ncclResult_t r = n->GetUniqueId(&id);
  if (r != 0) return nccl_check(nullptr, r, "ncclGetUniqueId");
  memcpy(id128, id.internal, 128);
  return ZOS_OK;
}

zos_status zos_comm_create(zos_ctx* ctx, const uint8_t* id128, uint32_t rank, uint32_t world, zos_comm** out) {
  if (!ctx || !id128 || !out || world == 0 || rank >= world) return ZOS_ERR_INVALID;
  *out = nullptr;
  Nccl* n = nccl();
  if (!n->lib) return zos::fail(ctx, ZOS_ERR_UNSUPPORTED, "NCCL unavailable: %s", n->why.c_str());
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  cudaSetDevice(ctx->device);
  ncclComm_t c = nullptr;
  zos_status st = nccl_check(ctx, n->CommInitRank(&c, (int)world, id, (int)rank), "ncclCommInitRank");
  if (st != ZOS_OK) return st;
  zos_comm* cm = new zos_comm();
  cm->ctx = ctx; cm->comm = c; cm->rank = rank; cm->world = world;
  *out = cm;
  return ZOS_OK;
}

void zos_comm_destroy(zos_comm* cm) {
  if (!cm) return;
  if (cm->comm && nccl()->lib) {
    cudaSetDevice(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    nccl()->CommDestroy(cm->comm);
  }
  delete cm;
}

int32_t zos_comm_nccl_version(void) {
  Nccl* n = nccl();
  int v = 0;
  if (!n->lib || n->GetVersion(&v) != 0) return 0;
  return v;
}

zos_status zos_gather_nccl(zos_comm* cm, const zos_buf* send, uint64_t send_off, zos_buf* recv, const uint64_t* recv_offsets,
                           const uint64_t* shard_bytes, int32_t root) {
  if (!cm || !shard_bytes) return ZOS_ERR_INVALID;
  zos_ctx* ctx = cm->ctx;
  Nccl* n = nccl();
  const uint32_t W = cm->world, me = cm->rank;
  if (root >= (int32_t)W) return zos::fail(ctx, ZOS_ERR_INVALID, "gather_nccl: root %d of %u ranks", root, W);
  const bool receiver = root < 0 || (uint32_t)root == me;
  const uint64_t mine = shard_bytes[me];
  if (mine && !in_buf(send, send_off, mine)) return zos::fail(ctx, ZOS_ERR_INVALID, "gather_nccl: send range out of bounds");
  if (receiver) {
    if (!recv || !recv_offsets) return zos::fail(ctx, ZOS_ERR_INVALID, "gather_nccl: a receiving rank needs recv and recv_offsets");
    for (uint32_t r = 0; r < W; r++)
      if (!in_buf(recv, recv_offsets[r], shard_bytes[r])) return zos::fail(ctx, ZOS_ERR_INVALID, "gather_nccl: shard %u does not fit recv", r);
  }
  cudaSetDevice(ctx->device);
  // equal shards, back to back, to everyone: the library's all-gather (ring / NVLS as NCCL decides)
  bool uniform = root < 0 && mine > 0;
  for (uint32_t r = 0; uniform && r < W; r++) uniform = shard_bytes[r] == mine && recv_offsets[r] == recv_offsets[0] + (uint64_t)r * mine;
  if (uniform)
    return nccl_check(ctx, n->AllGather((const uint8_t*)send->ptr + send_off, (uint8_t*)recv->ptr + recv_offsets[0], (size_t)mine, ncclUint8, cm->comm, ctx->stream),
                      "ncclAllGather");
  zos_status st = nccl_check(ctx, n->GroupStart(), "ncclGroupStart");
  if (st != ZOS_OK) return st;
  ncclResult_t r0 = 0;
  if (mine) {
    for (uint32_t r = 0; r < W && r0 == 0; r++) {
      if (!(root < 0 || (uint32_t)root == r) || r == me) continue;
      r0 = n->Send((const uint8_t*)send->ptr + send_off, (size_t)mine, ncclUint8, (int)r, cm->comm, ctx->stream);
    }
  }
  if (receiver) {
    for (uint32_t r = 0; r < W && r0 == 0; r++) {
      if (r == me || shard_bytes[r] == 0) continue;
      r0 = n->Recv((uint8_t*)recv->ptr + recv_offsets[r], (size_t)shard_bytes[r], ncclUint8, (int)r, cm->comm, ctx->stream);
    }
  }
  ncclResult_t r1 = n->GroupEnd();
  if (r0 != 0) return nccl_check(ctx, r0, "ncclSend/ncclRecv");
  if ((st = nccl_check(ctx, r1, "ncclGroupEnd")) != ZOS_OK) return st;
  if (receiver && mine)  // the rank's own shard: a device copy on the same stream
    return zos::check_cuda(ctx, cudaMemcpyAsync((uint8_t*)recv->ptr + recv_offsets[me], (const uint8_t*)send->ptr + send_off, mine, cudaMemcpyDeviceToDevice, ctx->stream),
                           "gather_nccl own shard");
  return ZOS_OK;
}

}  // extern "C"
