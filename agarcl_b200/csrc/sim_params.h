// sim_params.h — kernel parameter blocks shared by the host API (batch.cu) and the kernels.
#pragma once
#include <cstdint>

#include "../../include/agarcl_b200.h"
#include "device_math.cuh"

namespace ag {

constexpr int kWarpsPerCta = 4;       // one warp owns one game instance; a CTA is 4 independent warps
constexpr int kPremCap = 192;         // pellets_to_remove entries per tick (Engine.hpp:212)
constexpr int kVremCap = 32;          // viruses_to_remove entries per tick (Engine.hpp:213)
constexpr int kCandCap = 32;          // pellet candidates resolved in registers per cell
constexpr int kLaneCand = 8;          // pellet candidates a single lane resolves in the lane-per-player phase
constexpr int kSnapCap = 128;         // cells staged in shared memory by the players_collision pre-test
constexpr int kPairCap = 48;          // (eater, eaten) pairs per tick in players_collision
constexpr int kCellRefCap = 512;      // total live cells per instance handled by players_collision

struct SimParams {
  agarcl_layout L;
  Luts T;
  uint8_t* state;          // N blobs, L.stride apart
  const float* dxdy;       // [N*A*2] actions for this step
  const int32_t* act;      // [N*A]
  double* rewards;         // [N*A]
  uint8_t* dones;          // [N*A]
  float* before;           // [N*A] masses<float>() at step begin (BaseEnvironment.hpp:92)
  const float* replay;     // [N*cap_replay] or nullptr
  int32_t N;
  int32_t instance_base;
  int32_t n_ticks;         // ticks to run in this launch
  int32_t do_begin;        // apply actions + record `before`
  int32_t do_end;          // respawn / dones / rewards
  int32_t mode, reward_type, rng_mode;
  int32_t target_pellets, target_viruses;
  int32_t HG;              // spatial hash is HG x HG over the arena
  float W;                 // arena width == height
  float hash_scale;        // HG / W
  int32_t gw_pellet;       // reference pellet bucket grid width (bucket 510, Engine.hpp:962-965)
  int32_t gw_virus;        // reference virus bucket grid width (bucket 25, Engine.hpp:1207-1211)
  uint32_t smem_per_warp;  // bytes
};

struct ResetParams {
  agarcl_layout L;
  Luts T;
  uint8_t* state;
  const uint8_t* mask;     // [N] or nullptr (all)
  const uint64_t* seeds;   // [N]
  const float* replay;
  uint8_t* dones;
  int32_t N, instance_base, rng_mode, num_pellets, num_viruses;
  float W;
};

struct ObsParams {
  agarcl_layout L;
  const uint8_t* state;
  void* obs;               // [N*A][frames*C][G][G] int32 or int16
  int32_t N, G, C, frames, frame;  // frame: which frame slot this launch fills
  int32_t observe_cells, observe_others, observe_viruses, observe_pellets;
  int32_t obs_dtype;
  int32_t pre_respawn;     // 1: show players respawned at the end of the step as still dead
  float W;
};

}  // namespace ag
