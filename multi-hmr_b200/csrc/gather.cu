// Multi-GPU result exchange of the sharded batch path: images are independent units of
// Model.forward, every rank runs the whole engine on its contiguous image shard, and ONE all-gather of a
// compact per-rank record block returns every person to every rank.
//
//   block  = header (8 x int32) | capacity x record
//   header = { persons detected on this rank, persons packed (<= capacity), capacity, floats per record,
//              global index of the rank's first image, 0, 0, 0 }
//   record = [ global image index, score, loc(2), transl(3), transl_pelvis(3), rotvec(159), expression(10),
//              shape(nb), v3d(3V), j3d(381), j2d(254) ]   fp32 — the person dict of model.py:329-347
//
// The count travels in the header, so there is a single collective per step and no host round trip between
// the forward and the exchange.  mhmr_pack_records is ONE kernel; mhmr_allgather_records is ONE
// ncclAllGather on the caller's stream (NCCL is resolved at run time from the libnccl.so.2 the process
// already has loaded — PyTorch's — so that the library itself has no link-time dependency on it).
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "kernels.cuh"

using namespace mhmr;

struct Id128 { char bytes[128]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value

namespace {

constexpr int kHeaderWords = 8;

struct RecordSrc {
  const int* count;
  const int* det_b;
  const float *score, *loc, *transl, *transl_pelvis, *rotvec, *expression, *shape, *v3d, *j3d, *j2d;
};

// one CTA per (person slot, slice of the record); slots >= packed count are zero-filled so that the block is
// deterministic (bit-identical gathers for identical inputs)
__global__ void __launch_bounds__(256)
pack_records_kernel(RecordSrc s, int max_persons, int image_offset, int capacity, int nb, int V,
                    float* __restrict__ block) {
  const int R = 179 + nb + 3 * V + 381 + 254;
  const int P = min(*s.count, max_persons);
  const int packed = min(P, capacity);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < kHeaderWords) {
    int* hdr = reinterpret_cast<int*>(block);
    const int vals[kHeaderWords] = {*s.count, packed, capacity, R, image_offset, 0, 0, 0};
    hdr[threadIdx.x] = vals[threadIdx.x];
  }
  const int p = blockIdx.x;
  float* rec = block + kHeaderWords + static_cast<int64_t>(p) * R;
  const int chunk = (R + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * chunk, hi = min(lo + chunk, R);
  if (p >= packed) {
    for (int i = lo + threadIdx.x; i < hi; i += 256) rec[i] = 0.f;
    return;
  }
  const int o_shape = 179, o_v3d = 179 + nb, o_j3d = o_v3d + 3 * V, o_j2d = o_j3d + 381;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    float v;
    if (i >= o_v3d && i < o_j3d) v = s.v3d[static_cast<int64_t>(p) * 3 * V + (i - o_v3d)];
    else if (i == 0) v = static_cast<float>(s.det_b[p] + image_offset);
    else if (i == 1) v = s.score[p];
    else if (i < 4) v = s.loc[p * 2 + (i - 2)];
    else if (i < 7) v = s.transl[p * 3 + (i - 4)];
    else if (i < 10) v = s.transl_pelvis[p * 3 + (i - 7)];
    else if (i < 169) v = s.rotvec[p * 159 + (i - 10)];
    else if (i < 179) v = s.expression[p * 10 + (i - 169)];
    else if (i < o_v3d) v = s.shape[p * nb + (i - o_shape)];
    else if (i < o_j2d) v = s.j3d[p * 381 + (i - o_j3d)];
    else v = s.j2d[p * 254 + (i - o_j2d)];
    rec[i] = v;
  }
}

// ---- NCCL through dlopen ---------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // RTLD_NOLOAD first: reuse the copy PyTorch already mapped; then the loader's search path
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (lib == nullptr) lib = dlopen("libnccl.so.2", RTLD_NOW);
    if (lib == nullptr) return;
    api.lib = lib;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  });
  if (api.lib == nullptr || !api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) return nullptr;
  return &api;
}

int nccl_fail(NcclApi* a, int rc, const char* what) {
  set_last_error(std::string(what) + " failed: " + (a->GetErrorString ? a->GetErrorString(rc) : "NCCL error ") +
                 " (" + std::to_string(rc) + ")");
  return MHMR_ERR_CUDA;
}

}  // namespace

struct mhmr_comm {
  void* comm = nullptr;
  int world = 0, rank = 0;
};

extern "C" {

int mhmr_record_floats(int num_betas, int num_verts) { return 179 + num_betas + 3 * num_verts + 381 + 254; }

int64_t mhmr_record_block_bytes(int num_betas, int num_verts, int capacity) {
  return 4ll * (kHeaderWords + static_cast<int64_t>(capacity) * mhmr_record_floats(num_betas, num_verts));
}

int mhmr_pack_records(const mhmr_outputs* out, int max_persons, int num_betas, int num_verts, int image_offset,
                      int capacity, float* block, void* stream) {
  MHMR_REQUIRE(out != nullptr && block != nullptr, "null argument");
  MHMR_REQUIRE(capacity >= 1 && capacity <= max_persons, "capacity must be in [1, max_persons]");
  MHMR_REQUIRE(out->count && out->det_idx && out->det_score && out->loc && out->transl && out->transl_pelvis &&
                   out->rotvec && out->expression && out->shape && out->v3d && out->j3d && out->j2d,
               "a required output buffer is null");
  RecordSrc s{out->count, out->det_idx, out->det_score, out->loc, out->transl, out->transl_pelvis, out->rotvec,
              out->expression, out->shape, out->v3d, out->j3d, out->j2d};
  pack_records_kernel<<<dim3(capacity, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      s, max_persons, image_offset, capacity, num_betas, num_verts, block);
  MHMR_CUDA_CHECK(cudaGetLastError());
  return MHMR_OK;
}

int mhmr_nccl_unique_id(void* id128) {
  MHMR_REQUIRE(id128 != nullptr, "null argument");
  NcclApi* a = nccl_api();
  if (a == nullptr) { set_last_error("libnccl.so.2 is not available in this process"); return MHMR_ERR_UNSUPPORTED; }
  const int rc = a->GetUniqueId(id128);
  return rc == 0 ? MHMR_OK : nccl_fail(a, rc, "ncclGetUniqueId");
}

int mhmr_comm_create(const void* id128, int world, int rank, mhmr_comm** out) {
  MHMR_REQUIRE(id128 != nullptr && out != nullptr && world >= 1 && rank >= 0 && rank < world, "bad argument");
  NcclApi* a = nccl_api();
  if (a == nullptr) { set_last_error("libnccl.so.2 is not available in this process"); return MHMR_ERR_UNSUPPORTED; }
  Id128 id;
  memcpy(id.bytes, id128, 128);
  auto* c = new mhmr_comm();
  c->world = world;
  c->rank = rank;
  const int rc = a->CommInitRank(&c->comm, world, id, rank);
  if (rc != 0) { delete c; return nccl_fail(a, rc, "ncclCommInitRank"); }
  *out = c;
  return MHMR_OK;
}

int mhmr_comm_destroy(mhmr_comm* c) {
  if (c == nullptr) return MHMR_OK;
  NcclApi* a = nccl_api();
  if (a != nullptr && c->comm != nullptr) a->CommDestroy(c->comm);
  delete c;
  return MHMR_OK;
}

int mhmr_allgather_records(mhmr_comm* c, const void* block, void* all_blocks, int64_t block_bytes, void* stream) {
  MHMR_REQUIRE(c != nullptr && block != nullptr && all_blocks != nullptr && block_bytes > 0, "bad argument");
  NcclApi* a = nccl_api();
  if (a == nullptr) { set_last_error("libnccl.so.2 is not available in this process"); return MHMR_ERR_UNSUPPORTED; }
  const int rc = a->AllGather(block, all_blocks, static_cast<size_t>(block_bytes), /*ncclInt8*/ 0, c->comm,
                              static_cast<cudaStream_t>(stream));
  return rc == 0 ? MHMR_OK : nccl_fail(a, rc, "ncclAllGather");
}

}  // extern "C"
