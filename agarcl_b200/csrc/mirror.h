// mirror.h — host-resident observation mirror (mirror.cu), used by batch.cu behind
// agarcl_batch_mirror / agarcl_batch_sync_mirror / agarcl_batch_step_mirror.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ag {

struct HostMirror;

struct MirrorStats {
  uint64_t entries;      // non-zero elements of channels != 0 (mod C) moved by the last sync
  uint64_t dense_images; // images that took the dense copy (entry capacity exceeded or a non-separable mask)
  uint64_t d2h_bytes;    // bytes copied device -> host by the last sync
  uint64_t host_threads; // threads that expand the lists into the mirror
};

// n_img images of CH = frames*C channels of G x G elements; dtype: agarcl_obs_dtype.  nullptr + agarcl_set_error on failure.
HostMirror* mirror_create(int n_img, int CH, int C, int G, int dtype);
void mirror_destroy(HostMirror* m);
void* mirror_ptr(HostMirror* m);
// Makes the host mirror identical to the device observation `d_obs` (work enqueued on `s`, then synchronised;
// everything enqueued on `s` before the call is complete when it returns).
int mirror_sync(HostMirror* m, const void* d_obs, cudaStream_t s);
void mirror_stats(const HostMirror* m, MirrorStats* out);

}  // namespace ag
