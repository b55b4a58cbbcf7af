// mirror.h — host-resident observation mirror (mirror.cu), used by batch.cu behind
// agarcl_batch_mirror / agarcl_batch_sync_mirror / agarcl_batch_step_mirror.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "sim_params.h"

namespace ag {

struct HostMirror;

struct MirrorStats {
  uint64_t entries;      // non-zero elements of channels != 0 (mod C) moved by the last sync
  uint64_t dense_images; // images that took the dense copy (entry capacity exceeded or a non-separable mask)
  uint64_t d2h_bytes;    // bytes copied device -> host by the last sync
  uint64_t host_threads; // threads that expand the lists into the mirror
  uint64_t wait_us;      // last sync: time the calling thread waited for the device (chunk flags + chunk copies)
  uint64_t total_us;     // last sync: first chunk wait to mirror complete
};

// n_img = instances * agents images of CH = frames*C channels of G x G elements; dtype: agarcl_obs_dtype.
// nullptr + agarcl_set_error on failure.
HostMirror* mirror_create(int n_img, int agents, int CH, int C, int G, int dtype, bool lists_only = false);
void mirror_destroy(HostMirror* m);
void* mirror_ptr(HostMirror* m);
// Makes the host mirror identical to the device observation `d_obs`: k_pack enqueued on `s`, then synchronised;
// everything enqueued on `s` before the call is complete when it returns.
int mirror_sync(HostMirror* m, const void* d_obs, cudaStream_t s);
// The fused path: hand mirror_pack_out() to ONE k_step launch on `s` (SimParams::pk), then call mirror_collect(m, d_obs, s):
// chunks are fetched and expanded while the kernel runs; `s` is synchronised when it returns.
// mirror_collect returns when every chunk has been flagged and expanded: the mirror and rewards_out / dones_out (from the
// records, may be null) are complete; the kernel itself may still be retiring on `s`.
// mirror_stage_actions copies the step's actions into host-mapped staging the kernel reads directly (no H2D copy).
PackOut mirror_pack_out(HostMirror* m);
int mirror_collect(HostMirror* m, const void* d_obs, cudaStream_t s, double* rewards_out, uint8_t* dones_out);
void mirror_stage_actions(HostMirror* m, const float* dxdy, const int32_t* act, const float** d_dxdy, const int32_t** d_act);
void mirror_stats(const HostMirror* m, MirrorStats* out);
// The lists themselves as the observation (agarcl_batch_step_lists, include/agarcl_b200.h): a HostMirror created with
// lists_only has no dense tensor and no worker threads; mirror_collect_lists waits for the chunk flags of the launch that
// got mirror_pack_out() and fills rewards / dones from the records; mirror_lists_view describes the current lists;
// mirror_lists_expand decodes one image into a dense frame (1: the image overflowed its slot, take it from the device).
int mirror_collect_lists(HostMirror* m, cudaStream_t s, double* rewards_out, uint8_t* dones_out);
void mirror_lists_view(const HostMirror* m, agarcl_obs_lists* out);
int mirror_lists_expand(const HostMirror* m, int img, void* dense_out);

}  // namespace ag
