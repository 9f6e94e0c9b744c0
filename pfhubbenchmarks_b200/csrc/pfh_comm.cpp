// Multi-GPU plumbing for z-slab decomposition: ghost-plane exchange with grouped
// ncclSend/ncclRecv and fp64 sum all-reduce, over NVLink 5 / NVSwitch.  Replaces PETSc's
// VecScatter halo and MPI_Allreduce inside KSP/SNES.  NCCL is loaded with
// dlopen so that a single-GPU ctx needs no NCCL at all, and so that under torch the already
// loaded libnccl.so.2 is shared.
#include <dlfcn.h>

#include "pfh_internal.h"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[PFH_NCCL_ID_BYTES];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclDouble = 8, ncclSum = 0 };  // ncclFloat64 = 8 in nccl.h (NCCL 2.x)

struct Api {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;  // optional (NCCL >= 2.18)
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Api& api() {
  static Api a;
  if (a.h) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.h) break;
  }
  if (!a.h) throw PfhError(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define L(sym) \
  *(void**)(&a.sym) = dlsym(a.h, "nccl" #sym); \
  if (!a.sym) throw PfhError("libnccl: missing symbol nccl" #sym)
  L(GetUniqueId); L(CommInitRank); L(CommDestroy); L(AllReduce); L(Send); L(Recv); L(GroupStart); L(GroupEnd);
  L(GetErrorString);
#undef L
  *(void**)(&a.CommSplit) = dlsym(a.h, "ncclCommSplit");
  return a;
}

void ck(ncclResult_t r, const char* what) {
  if (r != 0) throw PfhError(std::string("NCCL error in ") + what + ": " + api().GetErrorString(r));
}
}  // namespace

struct Comm {
  ncclComm_t comm = nullptr;   // reductions (ctx stream)
  ncclComm_t halo = nullptr;   // ghost-plane exchanges: a duplicate of `comm` (ncclCommSplit), so that an exchange on
                               // comm_stream may run while a reduction is in flight on the ctx stream; == comm if unavailable
  int rank = 0, nranks = 1;
};

int comm_unique_id(void* out) {
  ncclUniqueId id;
  ck(api().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(out, &id, PFH_NCCL_ID_BYTES);
  return 0;
}

Comm* comm_create(const pfh_config& cfg, cudaStream_t) {
  if (cfg.nranks <= 1) return nullptr;
  PFH_CHECK(cfg.nccl_unique_id != nullptr, "nranks > 1 needs pfh_config.nccl_unique_id");
  Comm* c = new Comm;
  c->rank = cfg.rank;
  c->nranks = cfg.nranks;
  ncclUniqueId id;
  memcpy(&id, cfg.nccl_unique_id, PFH_NCCL_ID_BYTES);
  ck(api().CommInitRank(&c->comm, cfg.nranks, id, cfg.rank), "ncclCommInitRank");
  c->halo = c->comm;
  if (api().CommSplit != nullptr) {
    ncclComm_t dup = nullptr;
    if (api().CommSplit(c->comm, 0, cfg.rank, &dup, nullptr) == 0 && dup != nullptr) c->halo = dup;
  }
  return c;
}

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->halo && c->halo != c->comm) api().CommDestroy(c->halo);
  if (c->comm) api().CommDestroy(c->comm);
  delete c;
}

void comm_allreduce_sum(pfh_ctx* c, double* d_buf, int n) {
  if (!c->comm) return;
  ProfScope ps(c, T_ALLREDUCE);
  ck(api().AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, c->comm->comm, c->stream), "ncclAllReduce");
}

// Ghost planes: rank r owns planes [z_begin, z_end); it sends its first plane down to r-1 and its
// last plane up to r+1, and receives r-1's last plane into glo and r+1's first plane into ghi.
void comm_halo(pfh_ctx* c, const double* base, long long fstride, int nfields, double* glo, double* ghi, cudaStream_t stream) {
  if (!c->comm) return;
  cudaStream_t st = stream ? stream : c->stream;
  ProfScope ps(c, T_HALO, 0.0, st);
  const Comm* cm = c->comm;
  const long long np = c->nplane;
  const long long last = (long long)(c->m3.nzl - 1) * np;
  ck(api().GroupStart(), "ncclGroupStart");
  for (int f = 0; f < nfields; ++f) {
    const double* fb = base + f * fstride;
    if (cm->rank > 0) {
      ck(api().Send(fb, (size_t)np, ncclDouble, cm->rank - 1, cm->halo, st), "ncclSend");
      ck(api().Recv(glo + f * np, (size_t)np, ncclDouble, cm->rank - 1, cm->halo, st), "ncclRecv");
    }
    if (cm->rank < cm->nranks - 1) {
      ck(api().Send(fb + last, (size_t)np, ncclDouble, cm->rank + 1, cm->halo, st), "ncclSend");
      ck(api().Recv(ghi + f * np, (size_t)np, ncclDouble, cm->rank + 1, cm->halo, st), "ncclRecv");
    }
  }
  ck(api().GroupEnd(), "ncclGroupEnd");
}

void comm_halo_n(pfh_ctx* c, const double* base, int nplanes, double* glo, double* ghi, cudaStream_t stream) {
  if (!c->comm) return;
  cudaStream_t st = stream ? stream : c->stream;
  ProfScope ps(c, T_HALO, 0.0, st);
  const Comm* cm = c->comm;
  const size_t cnt = (size_t)nplanes * c->nplane;
  const long long last = (long long)(c->m3.nzl - nplanes) * c->nplane;
  ck(api().GroupStart(), "ncclGroupStart");
  if (cm->rank > 0) {
    ck(api().Send(base, cnt, ncclDouble, cm->rank - 1, cm->halo, st), "ncclSend");
    ck(api().Recv(glo, cnt, ncclDouble, cm->rank - 1, cm->halo, st), "ncclRecv");
  }
  if (cm->rank < cm->nranks - 1) {
    ck(api().Send(base + last, cnt, ncclDouble, cm->rank + 1, cm->halo, st), "ncclSend");
    ck(api().Recv(ghi, cnt, ncclDouble, cm->rank + 1, cm->halo, st), "ncclRecv");
  }
  ck(api().GroupEnd(), "ncclGroupEnd");
}

bool comm_halo_is_separate(const pfh_ctx* c) { return c->comm != nullptr && c->comm->halo != c->comm->comm; }
