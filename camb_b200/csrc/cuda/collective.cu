// collective.cu -- the two exchange steps of ONE cosmology sharded over the GPUs of a box, issued from C
// on the library's own stream: no host synchronisation between a kernel and the collective that follows it.
//
//   cb_gather_sources   after loop A: every rank has evolved its own k-modes (cb_modes.active); the spline along k
//                       (InitSourceInterpolation, cmbmain.f90:1249-1265) needs every source k, so the rows are exchanged
//                       with ONE ncclAllGather: pack this rank's rows (n_src x n_tau values per mode) -> all-gather ->
//                       scatter every rank's rows into LinearSrc(n_k, n_src, n_tau).  Each value crosses NVLink once and is
//                       copied, never added: the gathered array is the single-GPU array bit for bit.
//   cb_allreduce_cl     after loop C over this rank's slice of the integration wavenumbers: ONE ncclAllReduce(sum) of the
//                       partial iCl (n_l x 3|6 doubles).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch has already loaded when the host program uses
// torch.distributed, the system copy otherwise), so the library has no link-time dependency on it and single-GPU users
// never load it.  The communicator is created from a unique id the host program distributes (rank 0 calls
// cb_comm_unique_id, the id travels over whatever channel the host has -- MPI in CAMB's MPI builds, the torch.distributed
// store in the Python harness of this repo).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every call goes through the pointers below

#include <string>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool nccl_load(std::string* err) {
    if (g_nccl.handle) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) { *err = std::string("NCCL not available: ") + dlerror(); return false; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { *err = std::string("NCCL symbol missing: ") + name; dlclose(h); return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather") SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return true;
}

// rows of this rank: packed[a + n_pack * col] = src[idx[a] + n_k * col], col = source + n_src * time step
__global__ void cb_pack_rows_kernel(const double* __restrict__ src, const int* __restrict__ idx, int n_act, int n_pack, int n_k, int ncol,
                                    double* __restrict__ packed) {
    const int col = blockIdx.x;
    for (int a = threadIdx.x; a < n_pack; a += blockDim.x) packed[a + (size_t)n_pack * col] = a < n_act ? src[idx[a] + (size_t)n_k * col] : 0.0;
}

// all ranks' rows back into place: owner[k] = rank that evolved mode k, pos[k] = its position in that rank's list
__global__ void cb_unpack_rows_kernel(const double* __restrict__ gathered, const int* __restrict__ owner, const int* __restrict__ pos, int n_pack,
                                      int n_k, int ncol, double* __restrict__ src) {
    const int col = blockIdx.x;
    for (int k = threadIdx.x; k < n_k; k += blockDim.x)
        src[k + (size_t)n_k * col] = gathered[(size_t)owner[k] * n_pack * ncol + pos[k] + (size_t)n_pack * col];
}

}  // namespace

bool cb_nccl_unique_id(char* id128, std::string* err) {
    if (!nccl_load(err)) return false;
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { *err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return false; }
    for (int i = 0; i < NCCL_UNIQUE_ID_BYTES; i++) id128[i] = id.internal[i];
    return true;
}

bool cb_nccl_init(void** comm, int nranks, int rank, const char* id128, std::string* err) {
    if (!nccl_load(err)) return false;
    ncclUniqueId id;
    for (int i = 0; i < NCCL_UNIQUE_ID_BYTES; i++) id.internal[i] = id128[i];
    ncclComm_t c = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&c, nranks, id, rank);
    if (r != ncclSuccess) { *err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return false; }
    *comm = c;
    return true;
}

void cb_nccl_destroy(void* comm) {
    if (comm && g_nccl.handle) g_nccl.CommDestroy((ncclComm_t)comm);
}

// d_idx: this rank's modes (ascending), d_owner / d_pos: [n_k]; d_packed: n_pack * ncol, d_gathered: nranks * n_pack * ncol
bool cb_nccl_gather_rows(void* comm, int nranks, double* d_src, int n_k, int ncol, const int* d_idx, int n_act, int n_pack, const int* d_owner,
                         const int* d_pos, double* d_packed, double* d_gathered, cudaStream_t st, std::string* err) {
    cb_pack_rows_kernel<<<ncol, 128, 0, st>>>(d_src, d_idx, n_act, n_pack, n_k, ncol, d_packed);
    ncclResult_t r = g_nccl.AllGather(d_packed, d_gathered, (size_t)n_pack * ncol, ncclFloat64, (ncclComm_t)comm, st);
    if (r != ncclSuccess) { *err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return false; }
    cb_unpack_rows_kernel<<<ncol, 128, 0, st>>>(d_gathered, d_owner, d_pos, n_pack, n_k, ncol, d_src);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { *err = std::string("gather kernels: ") + cudaGetErrorString(e); return false; }
    return true;
}

bool cb_nccl_allreduce_sum(void* comm, double* d_buf, size_t n, cudaStream_t st, std::string* err) {
    ncclResult_t r = g_nccl.AllReduce(d_buf, d_buf, n, ncclFloat64, ncclSum, (ncclComm_t)comm, st);
    if (r != ncclSuccess) { *err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return false; }
    return true;
}
