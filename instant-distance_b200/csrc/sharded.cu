// sharded.cu — the index sharded by PointId range across the GPUs of one box (BASELINE config 5).
//
// One process per GPU.  Every rank owns an independent HNSW over its contiguous range of the input rows (own shuffle, own
// PointId space, own entry point); every query is searched on every shard with the same ef; then ONE ncclAllGather of the
// per-shard top-k — packed as u64 keys (canonical distance bits << 32 | global row id) by K1's epilogue — and a merge
// kernel that keeps the k smallest keys of the union, ties broken by the lower global id.  No other collective anywhere.
// NCCL is bound at run time (dlopen of libnccl.so.2: the system 2.27 or whichever copy the host application already loaded).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "internal.cuh"

namespace idb {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(a.handle, "ncclAllGather"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.handle, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
        return a;
    }();
    return api;
}

#define NCCL_TRY(expr)                                                                                                   \
    do {                                                                                                                 \
        ncclResult_t r__ = (expr);                                                                                       \
        if (r__ != ncclSuccess) return ::idb::fail(IDB_ERR_NCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl().GetErrorString(r__)); \
    } while (0)

// K4: merge of the G per-shard k-lists of one query (one warp per query): keep the k smallest keys of the union.
// Keys are unique (global ids are), so rank(key) = #{keys smaller} is a permutation.  The lists are NOT assumed sorted by
// the full key: a shard orders exact-distance ties by its local PointId, the merged order is by global id.
__global__ void merge_topk_kernel(const uint64_t* all_keys /* G x nq x k */, uint32_t G, uint64_t nq, uint32_t k,
                                  uint32_t* out_ids, float* out_dist, uint32_t* out_len) {
    extern __shared__ uint64_t sm_keys[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint64_t* keys = sm_keys + (size_t)warp * G * k;
    const uint32_t total = G * k;
    for (uint64_t q = (uint64_t)blockIdx.x * wpb + warp; q < nq; q += (uint64_t)gridDim.x * wpb) {
        for (uint32_t t = lane; t < total; t += 32) {
            const uint32_t g = t / k, j = t - g * k;
            keys[t] = all_keys[((size_t)g * nq + q) * k + j];
        }
        __syncwarp();
        uint32_t found = 0;
        for (uint32_t t = lane; t < total; t += 32) {
            const uint64_t key = keys[t];
            if (key == kKeyNone) continue;
            uint32_t rank = 0;
            for (uint32_t i = 0; i < total; ++i) rank += keys[i] < key ? 1u : 0u;
            if (rank < k) {
                out_ids[q * k + rank] = (uint32_t)key;
                if (out_dist) out_dist[q * k + rank] = __uint_as_float((uint32_t)(key >> 32));
            }
        }
        // number of real results = min(k, total non-empty keys); pad the tail
        uint32_t real = 0;
        for (uint32_t t = lane; t < total; t += 32) real += keys[t] != kKeyNone ? 1u : 0u;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) real += __shfl_xor_sync(kFullMask, real, off);
        found = min(real, k);
        for (uint32_t j = found + lane; j < k; j += 32) {
            out_ids[q * k + j] = kInvalid;
            if (out_dist) out_dist[q * k + j] = __int_as_float(0x7f800000);
        }
        if (out_len && lane == 0) out_len[q] = found;
        __syncwarp();
    }
}

}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

}  // namespace idb

using namespace idb;

extern "C" {

idb_status idb_comm_unique_id(void* out) {
    if (!out) return fail(IDB_ERR_INVALID_ARG, "out is null");
    if (!nccl().ok) return fail(IDB_ERR_NCCL, "libnccl.so.2 could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    static_assert(IDB_UNIQUE_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId id;
    NCCL_TRY(nccl().GetUniqueId(&id));
    std::memcpy(out, &id, sizeof(id));
    return IDB_OK;
}

idb_status idb_comm_create(const void* unique_id, int32_t rank, int32_t world, int32_t device, idb_comm** out) {
    if (!unique_id || !out) return fail(IDB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(IDB_ERR_INVALID_ARG, "rank %d / world %d invalid", rank, world);
    if (!nccl().ok) return fail(IDB_ERR_NCCL, "libnccl.so.2 could not be loaded");
    CUDA_TRY(cudaSetDevice(device));
    auto* c = new (std::nothrow) Comm();
    if (!c) return fail(IDB_ERR_OOM, "host allocation failed");
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = nccl().CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(IDB_ERR_NCCL, "ncclCommInitRank failed: %s", nccl().GetErrorString(r));
    }
    *out = reinterpret_cast<idb_comm*>(c);
    return IDB_OK;
}

void idb_comm_free(idb_comm* comm) {
    Comm* c = reinterpret_cast<Comm*>(comm);
    if (!c) return;
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

idb_status idb_index_set_id_map(idb_index* index, const uint32_t* global_ids) {
    if (!index) return fail(IDB_ERR_INVALID_ARG, "index is null");
    Index* ix = reinterpret_cast<Index*>(index);
    std::lock_guard<std::mutex> lk(ix->mu);
    CUDA_TRY(cudaSetDevice(ix->device));
    if (!global_ids) {
        cudaFree(ix->d_id_map);
        ix->d_id_map = nullptr;
        return IDB_OK;
    }
    if (ix->n == 0) return IDB_OK;
    if (!ix->d_id_map) CUDA_TRY(cudaMalloc(&ix->d_id_map, ix->n * 4));
    CUDA_TRY(cudaMemcpyAsync(ix->d_id_map, global_ids, ix->n * 4, cudaMemcpyHostToDevice, ix->stream));
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return IDB_OK;
}

}  // extern "C" (reopened below)

namespace idb {
idb_status search_device_keys(Index* ix, Lane& ln, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k, uint32_t* d_ids,
                              uint64_t* d_keys);  // api.cu

// Per-shard K1 (its epilogue packs (distance bits, global id) keys) -> ONE ncclAllGather -> merge kernel, all on lane 0's stream.
// The caller holds lanes[0].mu.
static idb_status sharded_search_locked(Index* ix, Comm* c, Lane& ln, const float* d_queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                        uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len) {
    CUDA_TRY(cudaSetDevice(ix->device));
    // merge kernel: `wpb` warps per CTA, each with world * k keys in shared memory; check the launch BEFORE the collective is enqueued
    int max_smem = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ix->device));
    const size_t per_warp = (size_t)c->world * k * 8;
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > (size_t)max_smem) wpb >>= 1;
    if (per_warp * wpb > (size_t)max_smem)
        return fail(IDB_ERR_UNSUPPORTED, "world * k = %llu keys per query do not fit the merge kernel's shared memory (%d bytes)",
                    (unsigned long long)c->world * k, max_smem);
    const size_t per = (size_t)nq * k;
    CUDA_TRY(ensure_u64(ln.keys_local, ln.keys_local_cap, per));
    CUDA_TRY(ensure_u64(ln.keys_all, ln.keys_all_cap, per * c->world));
    CUDA_TRY(ensure_u32(ln.ids, ln.ids_cap, per));
    idb_status st = search_device_keys(ix, ln, d_queries, nq, ef_search, k, ln.ids, ln.keys_local);
    if (st != IDB_OK) return st;
    NCCL_TRY(nccl().AllGather(ln.keys_local, ln.keys_all, per, ncclUint64, c->comm, ln.stream));
    const size_t smem = (size_t)wpb * per_warp;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<uint64_t>((nq + wpb - 1) / wpb, (uint64_t)ix->num_sms * 8);
    merge_topk_kernel<<<grid, wpb * 32, smem, ln.stream>>>(ln.keys_all, (uint32_t)c->world, nq, k, d_out_ids, d_out_dist, d_out_len);
    CUDA_TRY(cudaGetLastError());
    ln.last_launches += 2;  // all-gather + merge
    return IDB_OK;
}
}  // namespace idb

extern "C" {

// All device pointers; d_queries is nq x dim.  Collective: every rank of `comm` calls it with the same queries.
idb_status idb_sharded_search_batch_device(idb_index* index, idb_comm* comm, const float* d_queries, uint64_t nq, uint32_t ef_search,
                                           uint32_t k, uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len) {
    if (!index || !comm) return fail(IDB_ERR_INVALID_ARG, "index/comm is null");
    Index* ix = reinterpret_cast<Index*>(index);
    Comm* c = reinterpret_cast<Comm*>(comm);
    if (nq == 0) return IDB_OK;
    if (!d_queries || !d_out_ids || k == 0) return fail(IDB_ERR_INVALID_ARG, "bad argument");
    Lane& ln = ix->lanes[0];
    std::lock_guard<std::mutex> lk(ln.mu);
    ix->last_lane.store(0);
    return sharded_search_locked(ix, c, ln, d_queries, nq, ef_search, k, d_out_ids, d_out_dist, d_out_len);
}

idb_status idb_sharded_search_batch_f32(idb_index* index, idb_comm* comm, const float* queries, uint64_t nq, uint32_t ef_search, uint32_t k,
                                        uint32_t* out_ids, float* out_dist, uint32_t* out_len) {
    if (!index || !comm) return fail(IDB_ERR_INVALID_ARG, "index/comm is null");
    Index* ix = reinterpret_cast<Index*>(index);
    Comm* c = reinterpret_cast<Comm*>(comm);
    if (nq == 0) return IDB_OK;
    if (!queries || !out_ids || k == 0) return fail(IDB_ERR_INVALID_ARG, "bad argument");
    Lane& ln = ix->lanes[0];
    std::lock_guard<std::mutex> lk(ln.mu);  // held across staging, search and copy-back: the staging buffers belong to this call
    ix->last_lane.store(0);
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(ensure_f32(ln.q2, ln.q2_cap, nq * ix->dim));
    CUDA_TRY(ensure_u32(ln.ids2, ln.ids2_cap, nq * k));
    CUDA_TRY(ensure_f32(ln.dist, ln.dist_cap, nq * k));
    CUDA_TRY(ensure_u32(ln.len, ln.len_cap, nq));
    CUDA_TRY(cudaMemcpyAsync(ln.q2, queries, nq * ix->dim * 4, cudaMemcpyHostToDevice, ln.stream));
    idb_status st = sharded_search_locked(ix, c, ln, ln.q2, nq, ef_search, k, ln.ids2, ln.dist, ln.len);
    if (st != IDB_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(out_ids, ln.ids2, nq * k * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (out_dist) CUDA_TRY(cudaMemcpyAsync(out_dist, ln.dist, nq * k * 4, cudaMemcpyDeviceToHost, ln.stream));
    if (out_len) CUDA_TRY(cudaMemcpyAsync(out_len, ln.len, nq * 4, cudaMemcpyDeviceToHost, ln.stream));
    uint32_t ctrl[16] = {0};
    if (ln.ctrl && ln.last_nq) CUDA_TRY(cudaMemcpyAsync(ctrl, ln.ctrl, 64, cudaMemcpyDeviceToHost, ln.stream));
    CUDA_TRY(cudaStreamSynchronize(ln.stream));
    if (ctrl[12] != 0) return fail(IDB_ERR_CAPACITY, "%u queries overflowed an internal per-query structure on this shard", ctrl[12]);
    return IDB_OK;
}

}  // extern "C"
