// libb200amg.so — the C-ABI device engine declared in include/b200amg.h.
//
// Owns device copies of every level of an AMG hierarchy and runs the whole solve phase of
// AlgebraicMultigrid.jl (src/multilevel.jl:152-239, src/smoother.jl, src/preconditioner.jl:12-24)
// on one B200: the cycle is a static sequence of kernels captured once per cycle type into a
// CUDA graph and replayed per iteration.  There is no CPU fallback: without a device every
// compute entry point fails with B200AMG_ERR_NO_DEVICE.
#include "engine_base.h"

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges are no-ops unless a profiler is attached
#include <nccl.h>   // types and prototypes only: the library is dlopen'ed when a partition is requested

#include "kernels.cuh"
#include "stream.cuh"
#include "host_csr.h"
#include "partition.h"
#include "cluster_gs.cuh"
#include "dsm_gs.cuh"
#include "block_plan.h"
#include "block_gs.cuh"
#include "pass_plan.h"
#include "pass_gs.cuh"
#include "peer_halo.cuh"
#include "block_params.h"
#include "staging.h"

using namespace b200amg;

// ------------------------------------------------------------------------------------------
// NCCL, resolved at run time (a single-GPU process never needs libnccl).  RTLD_NOLOAD first: a host
// that already initialised torch.distributed has its NCCL loaded under the same soname.
// ------------------------------------------------------------------------------------------
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  if (api.ok) return api;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) throw AmgError{B200AMG_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror()};
#define B200AMG_NCCL_SYM(field, name)                                                         \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(lib, name));                          \
  if (!api.field) throw AmgError{B200AMG_ERR_NCCL, std::string("libnccl lacks ") + name};
  B200AMG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  B200AMG_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  B200AMG_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  B200AMG_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  B200AMG_NCCL_SYM(GroupStart, "ncclGroupStart")
  B200AMG_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  B200AMG_NCCL_SYM(Send, "ncclSend")
  B200AMG_NCCL_SYM(Recv, "ncclRecv")
  B200AMG_NCCL_SYM(AllReduce, "ncclAllReduce")
  B200AMG_NCCL_SYM(Broadcast, "ncclBroadcast")
  B200AMG_NCCL_SYM(AllGather, "ncclAllGather")
#undef B200AMG_NCCL_SYM
  api.ok = true;
  return api;
}
#define NCCL_OK(expr)                                                                                      \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess) {                                                                               \
      char _b[512];                                                                                        \
      snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(_r), __FILE__, __LINE__); \
      throw AmgError{B200AMG_ERR_NCCL, _b};                                                                \
    }                                                                                                      \
  } while (0)



// The engine is ONE translation unit cut into files by subject (the kernels are templates launched from here):
#include "engine_objects.cuh"       // device operators, smoother matrices (upload-time plans), levels, the handle
#include "engine_launch.cuh"        // kernel launchers and the per-level choice of the Gauss-Seidel sweep
#include "engine_cycle.cuh"         // __solve!, cycle graphs, phase timers
#include "engine_partition.cuh"     // row-partitioned levels: halo exchange, partitioned cycle
#include "engine_api_helpers.cuh"   // what the entry points share

extern "C" {

#include "abi_hierarchy.inc"        // create / add_level / set_coarse / set_partition / finalize / destroy
#include "abi_solve.inc"            // solve / cycle / precond / pcg / smoothers / single operators
#include "abi_diagnostics.inc"      // info, timing, options

}  // extern "C"
// (abi_spgemm.cu: the device Galerkin products; abi_plans.cu: the host-only plan entry points)
