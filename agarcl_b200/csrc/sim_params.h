// sim_params.h — kernel parameter blocks shared by the host API (batch.cu) and the kernels.
#pragma once
#include <cstdint>

#include "../../include/agarcl_b200.h"
#include "device_math.cuh"

namespace ag {

constexpr int kMaxWarpsPerCta = 16;   // one warp owns one game instance; ONE persistent CTA per SM of as many independent warps as the
                                      // shared memory holds (16 x 13.9 KB at 1000 pellets) and the register file allows (4 per SM
                                      // sub-partition x 128 registers per thread)
constexpr int kPremCap = 192;         // pellets_to_remove entries per tick (Engine.hpp:212)
constexpr int kVremCap = 32;          // viruses_to_remove entries per tick (Engine.hpp:213)
constexpr int kCandCap = 32;          // pellet candidates resolved in registers per cell
constexpr int kLaneCand = 8;          // pellet candidates a single lane resolves in the lane-per-player phase
constexpr int kZeroTileBytes = 2560;  // CTA-shared all-zero tile, source of the TMA bulk stores that clear the observation
constexpr int kColdCtxBytes = 96;    // per-warp slot of ColdCtx (sim_kernel.cu) + 16 B of the solver pool's counters
constexpr int kSnapCap = 96;          // cells staged in shared memory by the players_collision pre-test (12 bytes each: x, y, mass | player << 24)
constexpr int kPairCap = 48;          // (eater, eaten) pairs per tick in players_collision
constexpr int kCellRefCap = 256;      // total live cells per instance handled by players_collision

// byte offsets of one warp's shared-memory arrays (sim_shared.cuh)
struct SmemOff {
  uint32_t hcnt, htmp, hsorted, spel, mbar, cellref, rows, strip, hitq, snap, vcache, psum, pcell, pairs, reskeys, resorder, cand, prem, vrem, lprem, cold;
  uint32_t sweep_in_hash;  // the exact collision sweep's scratch lies over hsorted: running it invalidates the pellet hash
};

// Transfer lists of the host-resident observation mirror (mirror.cu), produced on the device either by the fused
// observation finish of k_step (no second pass over the dense tensor) or by k_pack.  The agent images are grouped
// into CHUNKS of `ipc` consecutive images; one chunk = one contiguous block of 32-bit words
//   [0] entries used in the chunk (k_pack only)   [1..3] pad
//   rec[ipc][rec_words]   per slot: count (entries of the image, or 0xFFFFFFFF: does not fit the scheme, copy the
//                         image densely), base (its first entry in the chunk's entry array), the index of the image
//                         the slot describes (k_step fills the slots in the order of its schedule, so that the chunks
//                         complete one after the other while it runs; k_pack: slot = image), then per frame the row
//                         bit mask and the column bit mask of the out-of-bounds channel (bit i: row i out of bounds),
//                         then the agent's reward (f64, two words) and done flag of the step (k_step only)
//   entries[cap_chunk]    uint2 (op << 29 | element offset inside the image, operand): the integer operation the
//                         host replays on its copy of the element (kPkSet .. kPkMax)
// k_step writes its blocks STRAIGHT INTO PINNED HOST MEMORY (`chunks` is then the device address of a host-mapped
// buffer): every image owns the fixed slot [li * slot, (li + 1) * slot) of its chunk's entry array -- slack costs
// nothing because only the bytes written cross PCIe -- and the warp writes its entries as coalesced 16-byte
// stores while it scatters.  The warp that finishes the last instance of a chunk raises the chunk's flag in
// host-mapped memory, so the host expands chunk k while the kernel is still stepping the instances of chunk k+1
// and nothing is left to copy when the kernel ends.  k_pack writes compact blocks in device memory (one cursor per
// chunk), which the host fetches with one copy per chunk.
struct PackOut {
  uint32_t* chunks;          // nullptr: no lists wanted
  uint32_t* cursor;          // [n_chunks] running entry count (k_pack; self-rewinding)
  uint32_t* done;            // [n_chunks] images finished (self-rewinding)
  volatile uint32_t* flags;  // [n_chunks] host-mapped: last step sequence number whose chunk is complete; may be nullptr
  uint32_t chunk_words;      // words between consecutive chunk blocks
  uint32_t ipc;              // slots (images) per chunk (a multiple of the agents per instance)
  uint32_t rec_words;        // 3 + frames * 2 * MW + 3
  uint32_t cap_chunk;        // entry capacity of a chunk
  uint32_t slot;             // k_step: entries per image slot (even; ipc * slot <= cap_chunk)
  uint32_t n_img;
  uint32_t seq;
  int32_t MW;                // 32-bit words per row (or column) mask
};
constexpr uint32_t kPackDense = 0xFFFFFFFFu;
// entry operations: x = v | x += v | x = x ? min(x, v) : v | x = max(x, v)   (GridObservation's at_least_ / total_mass_ /
// min / max channel rules, environment/envs/GridEnvironment.hpp:212-232); k_pack lists final values as kPkSet
constexpr uint32_t kPkSet = 0u, kPkAdd = 1u, kPkMinNz = 2u, kPkMax = 3u;
constexpr uint32_t kPkOffMask = 0x1FFFFFFFu;
__host__ __device__ inline uint32_t pk_off_rec(const PackOut&) { return 4u; }
__host__ __device__ inline uint32_t pk_off_entries(const PackOut& k) { return (4u + k.ipc * k.rec_words + 3u) & ~3u; }

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
  uint32_t* tickets;       // [2] persistent-grid work counter: next instance, warps that have left (k_step without alignment)
  uint32_t* cost;          // [N] cycles every instance worked in this step (nullptr: not wanted)
  const uint32_t* perm;    // [N] cost-sorted instance order of the previous step (nullptr: identity)
  uint32_t* sched;         // [2] or nullptr: [0] schedule of this launch (1: aligned as tick_barrier says, 0: free-running), set by k_order
                           // from [1], the instances that held a multi-cell player at the end of the previous launch
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
  float r_pellet;          // radius_conversion(PELLET_MASS = 1), the table's entry (core/utils.hpp:8-11)
  int32_t gw_pellet;       // reference pellet bucket grid width (bucket 510, Engine.hpp:962-965)
  int32_t gw_virus;        // reference virus bucket grid width (bucket 25, Engine.hpp:1207-1211)
  uint32_t smem_per_warp;  // bytes
  uint32_t tiles_bytes;    // CTA-wide tiles in front of the warps' carve-ups: the zero tile, then the all-ones tile (fused finish only)
  SmemOff so;
  // fused observation clear: the engine-tick kernel streams the zeros of channels 1..C-1 of every
  // agent frame (state independent, 7/8 of all bytes of the step) while it computes; k_obs then only
  // writes channel 0 and scatters the entities.  zero_vec_per_agent == 0 disables it.
  void* obs;                    // frame slot of agent 0 of instance 0
  uint32_t zero_vec_per_agent;  // 16-byte vectors to clear per agent frame
  uint32_t zero_skip_vec;       // vectors of channel 0 in front of them
  uint32_t agent_stride_vec;    // vectors between consecutive agents' frame slots
  // fused observation finish (one frame): after the last tick each warp also writes channel 0
  // and scatters the entities of its instance, so the step is ONE kernel.  obs_finish == 0: k_obs does it; 1: int32; 2: int16.
  int32_t obs_finish, obs_G, obs_C;
  int32_t zero_chunks;          // > 0: the clear is queued in this many pieces (4 per tick); 0: spread over the ticks
  int32_t observe_cells, observe_others, observe_viruses, observe_pellets;
  int32_t tick_barrier;         // instruction-fetch alignment (step_instance), bit mask of the CTA barriers of a tick: 1 tick start, 2 around the
                                // pooled pair solver, 4 before players_collision, 8 before move_foods, 16 before apply_removals; 0: free-running warps
  int32_t dyn_stripes;          // aligned schedule: 1 = stripes handed out by the ticket counter (default), 0 = the static round-robin pairing
  int32_t align_group;          // warps per alignment group (a divisor-free choice: the last group of a CTA may be smaller)
  PackOut pk;                   // with obs_finish only
  int32_t inst_first;           // this launch steps instances [inst_first, inst_first + N) of the batch
};

struct ResetParams {
  agarcl_layout L;
  Luts T;
  uint8_t* state;
  const uint8_t* mask;     // [N] or nullptr (all)
  const uint64_t* seeds;   // [N]
  const float* replay;
  uint8_t* fresh;          // [N] != 0: the instance was (re)seeded since its last reset -> its draw stream starts at 0; otherwise the
                           // stream goes on where the last episode left it, like the reference's engine RNG across reset()
                           // (BaseEnvironment.hpp:179-204 does not reseed); cleared by the kernel
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
  int32_t skip_zero;       // 1: channels 1..C-1 were already cleared by the engine-tick kernel
  const uint8_t* mask;     // [N] or nullptr: render only the instances with mask != 0 (masked reset)
  float W;
};

struct RamParams {
  agarcl_layout L;
  Luts T;
  const uint8_t* state;
  float* ram;              // [N][P][AGARCL_RAM_RECORD]
  int32_t N, G;
  int32_t pre_respawn;     // 1: players respawned at the end of the step are still dead (BaseEnvironment.hpp:96-101)
  int32_t pid_base;        // CloneInfo::owner is the player's pid: index + pid of player 0 (non-zero only in later episodes of strict_reference, Q3)
};

}  // namespace ag
