// comm.cu -- NCCL bound at run time; see comm.h.
#include "comm.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace gbc {

namespace {

// the slice of nccl.h this file needs (ABI of NCCL 2.x)
struct NcclUniqueId {
  char internal[128];
};
typedef struct ncclComm *NcclComm;
enum { kNcclSuccess = 0 };
enum { kNcclInt8 = 0, kNcclFloat64 = 8 };
enum { kNcclSum = 0 };

struct Api {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string err;
};

Api &api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      // a copy that the process already loaded (torch's bundled NCCL) wins over the system's
      a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
      if (!a.lib) a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) {
      a.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
      return;
    }
    auto sym = [&](const char *s) {
      void *p = dlsym(a.lib, s);
      if (!p && a.err.empty()) a.err = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return a;
}

int check(Api &a, int rc, const char *what, std::string &err) {
  if (rc == kNcclSuccess) return 0;
  err = std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error");
  return -1;
}

}  // namespace

struct Comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
};

int comm_unique_id(void *out128, std::string &err) {
  Api &a = api();
  if (!a.err.empty()) {
    err = a.err;
    return -1;
  }
  NcclUniqueId id;
  if (check(a, a.GetUniqueId(&id), "ncclGetUniqueId", err)) return -1;
  std::memcpy(out128, &id, sizeof(id));
  return 0;
}

Comm *comm_create(const void *id128, int rank, int world, std::string &err) {
  Api &a = api();
  if (!a.err.empty()) {
    err = a.err;
    return nullptr;
  }
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  Comm *c = new Comm;
  c->rank = rank;
  c->world = world;
  if (check(a, a.CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank", err)) {
    delete c;
    return nullptr;
  }
  return c;
}

void comm_destroy(Comm *c) {
  if (!c) return;
  if (c->comm) api().CommDestroy(c->comm);
  delete c;
}

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_world(const Comm *c) { return c ? c->world : 1; }

int comm_allreduce_sum(Comm *c, double *buf, size_t n, cudaStream_t st, std::string &err) {
  Api &a = api();
  return check(a, a.AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, c->comm, st), "ncclAllReduce",
               err);
}

int comm_allgather_inplace(Comm *c, void *buf, size_t bytesPerRank, cudaStream_t st,
                           std::string &err) {
  Api &a = api();
  const char *base = static_cast<const char *>(buf);
  return check(a,
               a.AllGather(base + (size_t)c->rank * bytesPerRank, buf, bytesPerRank, kNcclInt8,
                           c->comm, st),
               "ncclAllGather", err);
}

}  // namespace gbc
