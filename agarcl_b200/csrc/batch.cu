// batch.cu — host side of the C ABI (include/agarcl_b200.h): owns the device memory of a batch of
// N lockstep instances and enqueues the kernels.  Mirrors, for the batched path, what
// environment/bindings.cpp:99-135 binds on agario::env::GridEnvironment (ctor, seed,
// configure_observation, take_actions, step, get_state, dones, reset).
//
// There is NO CPU fallback: every entry point that computes needs a CUDA device and fails with
// AGARCL_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <random>
#include <unordered_map>
#include <vector>

#include "host_util.h"
#include "mirror.h"
#include "sim_params.h"
#include "sim_shared.cuh"

namespace ag {
cudaError_t launch_step(const SimParams& P, cudaStream_t stream);
cudaError_t selftest_std_sort(const float* ys, int n, uint16_t* order_out);
cudaError_t launch_order(const uint32_t* cost, uint32_t* perm, int N, uint32_t* sched, cudaStream_t stream);
cudaError_t launch_obs(const ObsParams& P, cudaStream_t stream);
cudaError_t launch_reset(const ResetParams& P, cudaStream_t stream);
cudaError_t launch_ram(const RamParams& P, cudaStream_t stream);
cudaError_t launch_flags(const uint8_t* state, const agarcl_layout& L, int N, uint32_t* out33, cudaStream_t stream);
}  // namespace ag

struct agarcl_batch {
  agarcl_cfg cfg;
  agarcl_layout L;
  // strict_reference, quirk Q3: the reference's player map and pid counter outlive a reset (GameState::clear keeps the bucket
  // array, Engine::reset does not rewind next_pid), so the iteration order of the players depends on how many times the
  // environment has been reset.  The same container, fed the same insert sequence, reset after reset.
  std::unordered_map<unsigned short, int> pmap;
  unsigned short next_pid = 0;
  int32_t pid_base = 0;  // pid of player index 0 in the current episode
  int N, A, G, C, frames;
  size_t obs_elems, obs_bytes;
  uint8_t* d_state = nullptr;
  void* d_obs = nullptr;
  float* d_ram = nullptr;  // structured observation records [N][P][AGARCL_RAM_RECORD], only with cfg.ram_obs
  double* d_rewards = nullptr;
  uint8_t* d_dones = nullptr;
  float* d_before = nullptr;
  float* d_dxdy = nullptr;
  int32_t* d_act = nullptr;
  const float* cur_dxdy = nullptr;  // what the next step reads (own buffers or caller's device pointers)
  const int32_t* cur_act = nullptr;
  float* d_replay = nullptr;
  uint64_t* d_seeds = nullptr;
  uint8_t* d_mask = nullptr;
  uint32_t* d_tickets = nullptr;  // k_step's ticket counter pair (self-rewinding)
  uint32_t* d_sched = nullptr;   // [2] schedule selector of k_step (SimParams::sched); AGARCL_AUTO_SCHEDULE=0 turns it off
  int auto_schedule = 1;
  uint32_t *d_cost = nullptr, *d_perm = nullptr;  // per-instance cost of the last step and the cost-sorted schedule made from it (k_order)
  bool perm_valid = false;
  int sort_schedule = 1;  // AGARCL_SORT_SCHEDULE=0 turns the cost-sorted schedule off (A/B timing)
  float *d_lut_radius = nullptr, *d_lut_speed = nullptr, *d_lut_split = nullptr;
  std::vector<uint64_t> seeds;
  // AGARCL_RNG_MT19937: the reference's generator per instance (GameState::rng, GameState.hpp:51), kept on the host; its
  // uniform_real_distribution<float> draws are fed to the device through d_replay used as a RING that refill_replay keeps
  // filled ahead of every instance's draw cursor (filled[i] = draws generated so far; the ring holds [filled - cap, filled))
  std::vector<std::mt19937_64> gens;
  std::vector<uint64_t> filled;
  int64_t draw_budget = 0;      // draws every instance is known to have ahead of its cursor
  int64_t max_draws_per_step = 0;
  uint8_t* d_fresh = nullptr;   // [N] instance (re)seeded since its last reset (ResetParams::fresh)
  uint32_t* d_flagbuf = nullptr;  // [33] agarcl_batch_flags scratch
  ag::Luts T;
  int HG;
  uint32_t smem_per_warp;
  ag::SmemOff so;
  bool was_reset = false;
  int fuse_clear = 2;  // 0: k_obs writes everything; 1: the engine-tick kernel clears channels 1..C-1 while ticking;
                       // 2: it also writes channel 0 and scatters the entities (one kernel per step).  AGARCL_FUSE_CLEAR overrides (A/B timing)
  int launches_last_step = 0;
  ag::HostMirror* mirror = nullptr;  // host-resident observation mirror (mirror.cu), created on first use
  ag::HostMirror* lists = nullptr;   // the lists-only twin behind agarcl_batch_step_lists (no dense tensor, no worker threads)
  uint64_t mirror_launch_us = 0, mirror_call_us = 0;  // last agarcl_batch_step_mirror: staging + launch, whole call
  // optional per-kernel timing: (start, after sim, after obs) event triples of steps not yet collected
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  double acc_sim_ms = 0.0, acc_obs_ms = 0.0;
  int acc_steps = 0;
};

static cudaEvent_t next_event(agarcl_batch* b) {
  if (b->ev_used == b->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    b->ev.push_back(e);
  }
  return b->ev[b->ev_used++];
}
static void collect_timing(agarcl_batch* b) {
  if (!b->ev_used) return;
  cudaEventSynchronize(b->ev[b->ev_used - 1]);
  for (size_t i = 0; i + 2 < b->ev_used; i += 3) {
    float a = 0.f, c = 0.f;
    cudaEventElapsedTime(&a, b->ev[i], b->ev[i + 1]);
    cudaEventElapsedTime(&c, b->ev[i + 1], b->ev[i + 2]);
    b->acc_sim_ms += a;
    b->acc_obs_ms += c;
    b->acc_steps++;
  }
  b->ev_used = 0;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return agarcl_set_error(AGARCL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));   \
  } while (0)

static void fill_sim_params(const agarcl_batch* b, ag::SimParams& P) {
  P.L = b->L;
  P.T = b->T;
  P.state = b->d_state;
  P.dxdy = b->cur_dxdy;
  P.act = b->cur_act;
  P.rewards = b->d_rewards;
  P.dones = b->d_dones;
  P.before = b->d_before;
  P.replay = b->d_replay;
  P.tickets = b->d_tickets;
  P.cost = nullptr;
  P.perm = nullptr;
  P.sched = nullptr;
  P.N = b->N;
  P.instance_base = b->cfg.instance_base;
  P.mode = b->cfg.mode_number;
  P.reward_type = b->cfg.reward_type;
  P.rng_mode = b->cfg.rng_mode;
  P.target_pellets = b->cfg.num_pellets;
  P.target_viruses = b->cfg.num_viruses;
  P.HG = b->HG;
  P.W = (float)b->cfg.arena_size;
  P.hash_scale = (float)b->HG / (float)b->cfg.arena_size;
  P.r_pellet = (float)std::sqrt((double)AGARCL_PELLET_MASS / 1.0 / M_PI);
  // initialize_pellet_grid / initialize_virus_grid: int((W + size - 1) / size) in fp32 (Engine.hpp:962-965,1207-1211)
  P.gw_pellet = (int)(((float)b->cfg.arena_size + 510.0f - 1.0f) / 510.0f);
  P.gw_virus = (int)(((float)b->cfg.arena_size + 25.0f - 1.0f) / 25.0f);
  P.smem_per_warp = b->smem_per_warp;
  P.tiles_bytes = ag::kZeroTileBytes;  // (+ the all-ones tile when the kernel also finishes the observation, fuse_obs_clear)
  P.so = b->so;
  P.obs = nullptr;
  P.zero_vec_per_agent = 0; P.zero_skip_vec = 0; P.agent_stride_vec = 0;
  P.obs_finish = 0; P.obs_G = b->G; P.obs_C = b->C;
  P.zero_chunks = 0;
  if (const char* e = std::getenv("AGARCL_ZERO_CHUNKS")) P.zero_chunks = std::atoi(e);
  P.dyn_stripes = 1;
  if (const char* e = std::getenv("AGARCL_DYNAMIC_STRIPES")) P.dyn_stripes = std::atoi(e);  // (A/B timing)
  P.tick_barrier = 6;  // bit mask of the alignment barriers of a tick (sim_kernel.cu, step_instance)
  if (const char* e = std::getenv("AGARCL_TICK_BARRIER")) P.tick_barrier = std::atoi(e);  // (A/B timing)
  P.align_group = ag::kMaxWarpsPerCta;  // the whole CTA (groups of 8 / 4 / 2 warps were slower: 1.78 / 1.97 / 2.16 vs 1.70 ms)
  P.observe_cells = b->cfg.observe_cells; P.observe_others = b->cfg.observe_others;
  P.observe_viruses = b->cfg.observe_viruses; P.observe_pellets = b->cfg.observe_pellets;
  std::memset(&P.pk, 0, sizeof(P.pk));
  P.inst_first = 0;
}

// Lets the engine-tick kernel clear channels 1..C-1 of frame slot `frame` (see SimParams); false when
// the planes are not 16-byte multiples.
static bool fuse_obs_clear(const agarcl_batch* b, ag::SimParams& P, int frame) {
  const size_t esz = b->cfg.obs_dtype == AGARCL_OBS_I16 ? 2 : 4;
  const size_t plane = (size_t)b->G * b->G * esz;
  if (plane % 16 != 0 || b->C < 2 || b->fuse_clear == 0) return false;
  P.obs = (uint8_t*)b->d_obs + (size_t)frame * b->C * plane;
  P.zero_skip_vec = (uint32_t)(plane / 16);
  P.zero_vec_per_agent = (uint32_t)((size_t)(b->C - 1) * plane / 16);
  P.agent_stride_vec = (uint32_t)((size_t)b->frames * b->C * plane / 16);
  // the whole observation in the engine-tick kernel: rows of whole 16-byte vectors, masks fit the scratch (1: int32, 2: int16)
  P.obs_finish = (b->fuse_clear >= 2 && ((size_t)b->G * esz) % 16 == 0 &&
                  (size_t)b->G * esz <= (size_t)ag::kZeroTileBytes &&  // one row of channel 0 = one bulk store from the ones tile
                  (size_t)b->G * esz <= (size_t)(b->so.vcache - b->so.cellref)) ? (esz == 2 ? 2 : 1) : 0;
  if (P.obs_finish) P.tiles_bytes = ag::kZeroTileBytes + (((uint32_t)b->G * (uint32_t)esz + 127u) & ~127u);
  return true;
}

static void fill_obs_params(const agarcl_batch* b, ag::ObsParams& P, int frame, int pre_respawn, int skip_zero = 0) {
  P.pre_respawn = pre_respawn;
  P.skip_zero = skip_zero;
  P.mask = nullptr;
  P.L = b->L;
  P.state = b->d_state;
  P.obs = b->d_obs;
  P.N = b->N;
  P.G = b->G;
  P.C = b->C;
  P.frames = b->frames;
  P.frame = frame;
  P.observe_cells = b->cfg.observe_cells;
  P.observe_others = b->cfg.observe_others;
  P.observe_viruses = b->cfg.observe_viruses;
  P.observe_pellets = b->cfg.observe_pellets;
  P.obs_dtype = b->cfg.obs_dtype;
  P.W = (float)b->cfg.arena_size;
}

// ---- AGARCL_RNG_MT19937 stream plumbing
static inline float mt_next(std::mt19937_64& g) {  // exactly what agarcl_mt19937_draws / Engine::random<float> draw (random.hpp:6-20)
  std::uniform_real_distribution<float> d(0.0f, 1.0f);
  return d(g);
}
// Generates draws [filled, upto) of instance i and stores them at their ring positions on the device.
static int mt_extend(agarcl_batch* b, int i, uint64_t upto) {
  const uint64_t cap = (uint64_t)b->L.cap_replay;
  if (upto <= b->filled[i]) return AGARCL_OK;
  uint64_t from = b->filled[i];
  if (upto - from > cap) {  // (only the last `cap` draws can live in the ring)
    for (uint64_t k = from; k < upto - cap; k++) (void)mt_next(b->gens[i]);
    from = upto - cap;
  }
  std::vector<float> buf((size_t)(upto - from));
  for (auto& v : buf) v = mt_next(b->gens[i]);
  float* row = b->d_replay + (size_t)i * cap;
  uint64_t k = from;
  while (k < upto) {
    const uint64_t pos = k % cap, n = std::min<uint64_t>(upto - k, cap - pos);
    if (cudaMemcpy(row + pos, buf.data() + (k - from), (size_t)n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
      return agarcl_set_error(AGARCL_ERR_CUDA, "replay ring upload failed");
    k += n;
  }
  b->filled[i] = upto;
  return AGARCL_OK;
}
// (Re)starts instance i's stream at draw index `at` from `seed` (Engine::seed, Engine.hpp:242-245) and fills the ring ahead of it.
static int mt_restart(agarcl_batch* b, int i, uint64_t seed, uint64_t at) {
  b->gens[i].seed((unsigned)seed);  // BaseEnvironment::seed(int) -> Engine::seed(unsigned)
  b->filled[i] = 0;
  for (uint64_t k = 0; k < at; k++) (void)mt_next(b->gens[i]);
  b->filled[i] = at;
  return mt_extend(b, i, at + (uint64_t)b->L.cap_replay);
}
// Called behind every launch that may draw (step, reset): when the instances may have come within one step's worth of
// draws of the end of what the ring holds, read the cursors back and fill the ring ahead of them again.
static int refill_replay(agarcl_batch* b, cudaStream_t s) {
  if (b->cfg.rng_mode != AGARCL_RNG_MT19937) return AGARCL_OK;
  b->draw_budget -= b->max_draws_per_step;
  if (b->draw_budget >= b->max_draws_per_step) return AGARCL_OK;
  std::vector<uint32_t> cur((size_t)b->N);
  if (cudaMemcpy2DAsync(cur.data(), sizeof(uint32_t), b->d_state + b->L.off_hdr + offsetof(agarcl_inst_hdr, rng_cursor), b->L.stride,
                        sizeof(uint32_t), (size_t)b->N, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess)
    return agarcl_set_error(AGARCL_ERR_CUDA, "draw cursor read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
  int64_t budget = b->L.cap_replay;
  for (int i = 0; i < b->N; i++) {
    // the 32-bit device cursor counts modulo 2^32; the stream position is the one nearest below `filled`
    uint64_t c = (b->filled[i] & ~0xffffffffull) | cur[i];
    if (c > b->filled[i]) c -= 0x100000000ull;
    const uint64_t ahead = b->filled[i] - c;
    if ((int64_t)ahead < (int64_t)b->L.cap_replay / 2 || (int64_t)ahead < 2 * b->max_draws_per_step) {
      int rc = mt_extend(b, i, c + (uint64_t)b->L.cap_replay);
      if (rc) return rc;
    }
    budget = std::min<int64_t>(budget, (int64_t)(b->filled[i] - c));
  }
  b->draw_budget = budget;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_destroy(agarcl_batch* b) {
  if (!b) return AGARCL_OK;
  ag::mirror_destroy(b->mirror);
  ag::mirror_destroy(b->lists);
  cudaFree(b->d_ram);
  cudaFree(b->d_state); cudaFree(b->d_obs); cudaFree(b->d_rewards); cudaFree(b->d_dones); cudaFree(b->d_before);
  cudaFree(b->d_dxdy); cudaFree(b->d_act); cudaFree(b->d_replay); cudaFree(b->d_seeds); cudaFree(b->d_mask); cudaFree(b->d_tickets); cudaFree(b->d_cost); cudaFree(b->d_perm);
  cudaFree(b->d_fresh); cudaFree(b->d_flagbuf); cudaFree(b->d_sched);
  cudaFree(b->d_lut_radius); cudaFree(b->d_lut_speed); cudaFree(b->d_lut_split);
  for (cudaEvent_t e : b->ev) cudaEventDestroy(e);
  delete b;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_create(const agarcl_cfg* cfg, agarcl_batch** out) {
  if (!cfg || !out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_instances < 1) return agarcl_set_error(AGARCL_ERR_INVALID, "n_instances must be >= 1");
  agarcl_layout L;
  int rc = agarcl_make_layout(cfg, &L);
  if (rc != AGARCL_OK) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return agarcl_set_error(AGARCL_ERR_CUDA, "no CUDA device: agarcl_b200 has no CPU path (%s)", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return agarcl_set_error(AGARCL_ERR_INVALID, "device %d out of range", cfg->device);
  CK(cudaSetDevice(cfg->device));
  agarcl_batch* b = new (std::nothrow) agarcl_batch();
  if (!b) return agarcl_set_error(AGARCL_ERR_NOMEM, "out of host memory");
  b->cfg = *cfg;
  b->L = L;
  b->N = cfg->n_instances;
  b->A = L.A;
  b->G = cfg->grid_size;
  b->C = L.obs_channels;
  b->frames = cfg->num_frames;
  if (const char* e = std::getenv("AGARCL_FUSE_CLEAR")) b->fuse_clear = std::atoi(e);
  if (const char* e = std::getenv("AGARCL_SORT_SCHEDULE")) b->sort_schedule = std::atoi(e);
  b->obs_elems = (size_t)b->N * b->A * b->frames * b->C * b->G * b->G;
  b->obs_bytes = b->obs_elems * (cfg->obs_dtype == AGARCL_OBS_I16 ? 2 : 4);
  // spatial hash resolution: about 3 pellets per hash cell, 4..64 cells per side
  double per_cell = 3.0;
  if (const char* e = std::getenv("AGARCL_HASH_PER_CELL")) per_cell = std::atof(e);  // (A/B timing)
  int hg = (int)std::floor(std::sqrt((double)L.cap_pellets / per_cell));
  b->HG = hg < 4 ? 4 : (hg > 64 ? 64 : hg);
  b->smem_per_warp = ag::make_smem_offsets(L, b->HG, b->so);
  if ((size_t)b->smem_per_warp + 2 * ag::kZeroTileBytes > 227 * 1024) {  // (room for one warp with both tiles at full size)
    delete b;
    return agarcl_set_error(AGARCL_ERR_INVALID, "configuration needs %u B of shared memory per instance (too many pellets/viruses)", b->smem_per_warp);
  }
#define ALLOC(ptr, bytes)                                                                           \
  do {                                                                                              \
    cudaError_t e__ = cudaMalloc((void**)&(ptr), (bytes));                                          \
    if (e__ != cudaSuccess) {                                                                       \
      agarcl_batch_destroy(b);                                                                      \
      return agarcl_set_error(AGARCL_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e__)); \
    }                                                                                               \
  } while (0)
  const size_t NA = (size_t)b->N * b->A;
  ALLOC(b->d_state, (size_t)b->N * L.stride);
  ALLOC(b->d_obs, b->obs_bytes);
  ALLOC(b->d_rewards, NA * sizeof(double));
  ALLOC(b->d_dones, NA);
  ALLOC(b->d_before, NA * sizeof(float));
  ALLOC(b->d_dxdy, NA * 2 * sizeof(float));
  ALLOC(b->d_act, NA * sizeof(int32_t));
  ALLOC(b->d_seeds, (size_t)b->N * sizeof(uint64_t));
  ALLOC(b->d_mask, (size_t)b->N);
  ALLOC(b->d_tickets, 2 * sizeof(uint32_t));
  ALLOC(b->d_cost, (size_t)b->N * sizeof(uint32_t));
  ALLOC(b->d_perm, (size_t)b->N * sizeof(uint32_t));
  cudaMemset(b->d_tickets, 0, 2 * sizeof(uint32_t));
  ALLOC(b->d_sched, 2 * sizeof(uint32_t));
  { const uint32_t init[2] = {1u, 0u}; cudaMemcpy(b->d_sched, init, sizeof(init), cudaMemcpyHostToDevice); }
  if (const char* e = std::getenv("AGARCL_AUTO_SCHEDULE")) b->auto_schedule = std::atoi(e);
  ALLOC(b->d_fresh, (size_t)b->N);
  cudaMemset(b->d_fresh, 1, (size_t)b->N);
  ALLOC(b->d_flagbuf, 33 * sizeof(uint32_t));
  if (L.cap_replay > 0) ALLOC(b->d_replay, (size_t)b->N * L.cap_replay * sizeof(float));
  if (cfg->ram_obs) {
    ALLOC(b->d_ram, (size_t)b->N * L.P * AGARCL_RAM_RECORD * sizeof(float));
    cudaMemset(b->d_ram, 0, (size_t)b->N * L.P * AGARCL_RAM_RECORD * sizeof(float));
  }
  ALLOC(b->d_lut_radius, AGARCL_LUT_SIZE * sizeof(float));
  ALLOC(b->d_lut_speed, AGARCL_LUT_SIZE * sizeof(float));
  ALLOC(b->d_lut_split, AGARCL_LUT_SIZE * sizeof(float));
#undef ALLOC
  cudaMemset(b->d_state, 0, (size_t)b->N * L.stride);
  cudaMemset(b->d_obs, 0, b->obs_bytes);
  cudaMemset(b->d_rewards, 0, NA * sizeof(double));
  cudaMemset(b->d_dones, 0, NA);
  cudaMemset(b->d_dxdy, 0, NA * 2 * sizeof(float));
  cudaMemset(b->d_act, 0, NA * sizeof(int32_t));
  if (b->d_replay) cudaMemset(b->d_replay, 0, (size_t)b->N * L.cap_replay * sizeof(float));
  b->cur_dxdy = b->d_dxdy;
  b->cur_act = b->d_act;
  // lookup tables of functions of an integer mass, computed with the host libm the reference uses
  {
    std::vector<float> rad(AGARCL_LUT_SIZE), spd(AGARCL_LUT_SIZE), spl(AGARCL_LUT_SIZE);
    for (uint32_t m = 0; m < AGARCL_LUT_SIZE; m++) {
      rad[m] = (float)std::sqrt((double)m / 1.0 / M_PI);                  // radius_conversion, core/utils.hpp:8-11
      spd[m] = (float)(300 / std::pow((double)m, 0.439));                 // max_speed, Engine.hpp:1300-1302
      double v = 3 * std::pow((double)spd[m], 1.2);                       // split_speed, Engine.hpp:1296-1298
      v = (130.0 < v) ? 130.0 : v;
      v = (v < 20.0) ? 20.0 : v;
      spl[m] = (float)v;
    }
    cudaMemcpy(b->d_lut_radius, rad.data(), rad.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(b->d_lut_speed, spd.data(), spd.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(b->d_lut_split, spl.data(), spl.size() * 4, cudaMemcpyHostToDevice);
    b->T.radius = b->d_lut_radius;
    b->T.max_speed = b->d_lut_speed;
    b->T.split_speed = b->d_lut_split;
    b->T.anti_team[0] = 1.0f;
    for (int n = 1; n <= AGARCL_VET_CAP; n++) b->T.anti_team[n] = (float)std::pow(1.1, (double)(n - 1));  // Engine.hpp:567
  }
  b->seeds.resize(b->N);
  for (int i = 0; i < b->N; i++) b->seeds[i] = (uint64_t)(cfg->instance_base + i);
  b->max_draws_per_step = 2 * ((int64_t)L.cap_pellets + L.cap_viruses + L.P);  // regen deficits + respawns of one step; a whole reset
  if (cfg->rng_mode == AGARCL_RNG_MT19937) {
    // an UNSEEDED reference environment draws from std::random_device (GameState.hpp:59): so does an unseeded batch
    if ((int64_t)L.cap_replay <= b->max_draws_per_step) {
      agarcl_batch_destroy(b);
      return agarcl_set_error(AGARCL_ERR_INVALID, "cap_replay %d cannot hold the draws of one step (%lld): raise cfg.cap_replay", L.cap_replay,
                              (long long)b->max_draws_per_step);
    }
    b->gens.resize(b->N);
    b->filled.assign(b->N, 0);
    std::random_device rd;
    for (int i = 0; i < b->N; i++) {
      uint64_t s0;
      try { s0 = rd(); } catch (...) { s0 = (uint64_t)(cfg->instance_base + i) * 0x9E3779B97F4A7C15ull + 1u; }
      b->seeds[i] = s0 & 0xffffffffull;
      int rc2 = mt_restart(b, i, b->seeds[i], 0);
      if (rc2) { agarcl_batch_destroy(b); return rc2; }
    }
    b->draw_budget = L.cap_replay;
  }
  cudaMemcpy(b->d_seeds, b->seeds.data(), b->N * sizeof(uint64_t), cudaMemcpyHostToDevice);
  CK(cudaDeviceSynchronize());
  *out = b;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_get_layout(const agarcl_batch* b, agarcl_layout* out) {
  if (!b || !out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  *out = b->L;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_seed(agarcl_batch* b, const uint64_t* seeds) {
  if (!b || !seeds) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  CK(cudaSetDevice(b->cfg.device));
  for (int i = 0; i < b->N; i++) b->seeds[i] = seeds[i];
  CK(cudaMemcpy(b->d_seeds, b->seeds.data(), b->N * sizeof(uint64_t), cudaMemcpyHostToDevice));
  if (b->cfg.rng_mode == AGARCL_RNG_MT19937) {
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < b->N; i++) {
      int rc = mt_restart(b, i, seeds[i], 0);
      if (rc) return rc;
    }
    b->draw_budget = b->L.cap_replay;
  }
  CK(cudaMemset(b->d_fresh, 1, (size_t)b->N));  // seed() restarts the stream: the next reset draws from index 0
  return AGARCL_OK;
}

extern "C" int agarcl_batch_set_replay(agarcl_batch* b, int32_t instance, const float* draws, int32_t n) {
  if (!b || !draws) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (instance < 0 || instance >= b->N) return agarcl_set_error(AGARCL_ERR_INVALID, "instance out of range");
  if (!b->d_replay) return agarcl_set_error(AGARCL_ERR_STATE, "batch was not created with a replay rng_mode");
  if (n > b->L.cap_replay) n = b->L.cap_replay;
  CK(cudaSetDevice(b->cfg.device));
  CK(cudaMemcpy(b->d_replay + (size_t)instance * b->L.cap_replay, draws, (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
  CK(cudaMemset(b->d_fresh + instance, 1, 1));  // a new stream starts at its first draw
  return AGARCL_OK;
}

static int render_frame(agarcl_batch* b, int frame, cudaStream_t s, int pre_respawn, int skip_zero = 0,
                        const uint8_t* d_mask = nullptr) {
  ag::ObsParams P;
  fill_obs_params(b, P, frame, pre_respawn, skip_zero);
  P.mask = d_mask;
  CK(ag::launch_obs(P, s));
  return AGARCL_OK;
}

static int render_ram(agarcl_batch* b, cudaStream_t s, int pre_respawn) {
  ag::RamParams P;
  P.L = b->L;
  P.T = b->T;
  P.state = b->d_state;
  P.ram = b->d_ram;
  P.N = b->N;
  P.G = b->G;
  P.pre_respawn = pre_respawn;
  P.pid_base = b->pid_base;
  CK(ag::launch_ram(P, s));
  return AGARCL_OK;
}

extern "C" int agarcl_batch_ram(agarcl_batch* b, float** dev_ptr, int64_t shape[3]) {
  if (!b || !dev_ptr) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (!b->d_ram) return agarcl_set_error(AGARCL_ERR_STATE, "batch was created without cfg.ram_obs");
  *dev_ptr = b->d_ram;
  if (shape) { shape[0] = b->N; shape[1] = b->L.P; shape[2] = AGARCL_RAM_RECORD; }
  return AGARCL_OK;
}

extern "C" int agarcl_batch_ram_host(agarcl_batch* b, float* out) {
  if (!b || !out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (!b->d_ram) return agarcl_set_error(AGARCL_ERR_STATE, "batch was created without cfg.ram_obs");
  CK(cudaSetDevice(b->cfg.device));
  CK(cudaMemcpy(out, b->d_ram, (size_t)b->N * b->L.P * AGARCL_RAM_RECORD * sizeof(float), cudaMemcpyDeviceToHost));
  return AGARCL_OK;
}

extern "C" int agarcl_batch_render_ram(agarcl_batch* b, void* stream) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  if (!b->d_ram) return agarcl_set_error(AGARCL_ERR_STATE, "batch was created without cfg.ram_obs");
  CK(cudaSetDevice(b->cfg.device));
  return render_ram(b, (cudaStream_t)stream, 0);
}

extern "C" int agarcl_batch_render(agarcl_batch* b, void* stream) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  CK(cudaSetDevice(b->cfg.device));
  return render_frame(b, b->frames - 1, (cudaStream_t)stream, 0);
}

extern "C" int agarcl_batch_reset(agarcl_batch* b, const uint8_t* mask, void* stream) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  CK(cudaSetDevice(b->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  ag::ResetParams P;
  P.L = b->L;
  P.T = b->T;
  P.state = b->d_state;
  P.mask = nullptr;
  if (mask) {
    CK(cudaMemcpyAsync(b->d_mask, mask, b->N, cudaMemcpyHostToDevice, s));
    P.mask = b->d_mask;
  }
  P.seeds = b->d_seeds;
  P.replay = b->d_replay;
  P.fresh = b->d_fresh;
  P.dones = b->d_dones;
  P.N = b->N;
  P.instance_base = b->cfg.instance_base;
  P.rng_mode = b->cfg.rng_mode;
  P.num_pellets = b->cfg.num_pellets;
  P.num_viruses = b->cfg.num_viruses;
  P.W = (float)b->cfg.arena_size;
  if (b->cfg.strict_reference && !mask) {
    // BaseEnvironment::reset of every instance (BaseEnvironment.hpp:179-197): players.clear(), then add_player for the agents and
    // the bots with the pids that follow the last episode's; a masked reset (one instance of a vector env) keeps the batch's order
    b->pmap.clear();
    b->pid_base = (int32_t)b->next_pid;
    for (int p = 0; p < b->L.P; p++) b->pmap.insert(std::make_pair(b->next_pid++, p));
    int k = 0;
    for (auto& kv : b->pmap) b->L.order[k++] = kv.second;
    P.L = b->L;
  }
  CK(ag::launch_reset(P, s));
  b->was_reset = true;
  { int rc = refill_replay(b, s); if (rc) return rc; }
  if (b->d_ram) {  // GoBiggerEnvironment::reset ends with observation.clear() (GoBiggerEnvironment.hpp:698-702)
    const size_t per = (size_t)b->L.P * AGARCL_RAM_RECORD * sizeof(float);
    if (!mask) CK(cudaMemsetAsync(b->d_ram, 0, per * b->N, s));
    else
      for (int i = 0; i < b->N; i++)
        if (mask[i]) CK(cudaMemsetAsync((uint8_t*)b->d_ram + per * i, 0, per, s));
  }
  // the reference ends reset() with _partial_observation (BaseEnvironment.hpp:202-203)
  // (a masked reset touches only the observations of the instances it resets)
  if (b->cfg.ram_obs == 2) return AGARCL_OK;  // no grid observation in this mode
  const uint8_t* dm = mask ? b->d_mask : nullptr;
  auto clear_obs = [&]() -> cudaError_t {
    if (!mask) return cudaMemsetAsync(b->d_obs, 0, b->obs_bytes, s);
    const size_t per = b->obs_bytes / (size_t)b->N;
    for (int i = 0; i < b->N; i++)
      if (mask[i]) {
        cudaError_t e = cudaMemsetAsync((uint8_t*)b->d_obs + per * i, 0, per, s);
        if (e != cudaSuccess) return e;
      }
    return cudaSuccess;
  };
  if (b->cfg.strict_reference) {
    CK(clear_obs());
    int frame = 0 - (b->cfg.ticks_per_step - b->frames);
    if (frame >= 0) return render_frame(b, frame, s, 0, 0, dm);
    return AGARCL_OK;
  }
  if (b->frames > 1) CK(clear_obs());
  return render_frame(b, b->frames - 1, s, 0, 0, dm);
}

extern "C" int agarcl_batch_set_actions(agarcl_batch* b, const float* dxdy, const int32_t* act, int on_device, void* stream) {
  if (!b || !dxdy || !act) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  CK(cudaSetDevice(b->cfg.device));
  const size_t NA = (size_t)b->N * b->A;
  if (on_device) {
    b->cur_dxdy = dxdy;
    b->cur_act = act;
  } else {
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(b->d_dxdy, dxdy, NA * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b->d_act, act, NA * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    b->cur_dxdy = b->d_dxdy;
    b->cur_act = b->d_act;
  }
  return AGARCL_OK;
}

// One env-step of every instance.  `want_lists`: the caller is agarcl_batch_step_mirror; when the step is the single
// fused kernel, it also leaves the host mirror's transfer lists (*lists_made = true) and the k_pack pass is not needed.
static int step_impl(agarcl_batch* b, cudaStream_t s, bool want_lists, bool* lists_made, ag::HostMirror* target = nullptr) {
  if (!target) target = b->mirror;
  if (lists_made) *lists_made = false;
  if (!b->was_reset) return agarcl_set_error(AGARCL_ERR_STATE, "step() before reset()");
  CK(cudaSetDevice(b->cfg.device));
  ag::SimParams P;
  fill_sim_params(b, P);
  const int tps = b->cfg.ticks_per_step;
  int launches = 0;
  if (b->cfg.strict_reference) {
    // quirk Q11: one _partial_observation(agent, 0) after all ticks, frame_index = -(tps - num_frames)
    P.n_ticks = tps; P.do_begin = 1; P.do_end = 1;
    CK(cudaMemsetAsync(b->d_obs, 0, b->obs_bytes, s));  // _step_hook: clear_data
    CK(ag::launch_step(P, s)); launches++;
    int frame = 0 - (tps - b->frames);
    if (frame >= 0) { int rc = render_frame(b, frame, s, 1); if (rc) return rc; launches++; }
  } else if (b->cfg.ram_obs == 2) {
    // structured observation only (agario-ram-v0): no grid frame is rendered at all
    P.n_ticks = tps; P.do_begin = 1; P.do_end = 1;
    const bool sorted = P.tick_barrier && b->sort_schedule;
    if (sorted) { P.cost = b->d_cost; P.perm = b->perm_valid ? b->d_perm : nullptr; P.sched = b->auto_schedule ? b->d_sched : nullptr; }
    if (b->timing) { if (b->ev_used >= 3 * 2048) collect_timing(b); CK(cudaEventRecord(next_event(b), s)); }
    CK(ag::launch_step(P, s)); launches++;
    if (b->timing) CK(cudaEventRecord(next_event(b), s));
    if (sorted) { CK(ag::launch_order(b->d_cost, b->d_perm, b->N, P.sched, s)); b->perm_valid = true; launches++; }
  } else if (b->frames == 1) {
    P.n_ticks = tps; P.do_begin = 1; P.do_end = 1;
    const int fused = fuse_obs_clear(b, P, 0) ? 1 : 0;
    if (want_lists && P.obs_finish && target && 6 + 2 * ((b->G + 31) / 32) <= 32) {  // (the image's record is one warp store)
      P.pk = ag::mirror_pack_out(target);
      *lists_made = true;
    }
    const bool sorted = P.tick_barrier && b->sort_schedule;
    if (sorted) { P.cost = b->d_cost; P.perm = b->perm_valid ? b->d_perm : nullptr; P.sched = b->auto_schedule ? b->d_sched : nullptr; }
    if (b->timing) { if (b->ev_used >= 3 * 2048) collect_timing(b); CK(cudaEventRecord(next_event(b), s)); }
    CK(ag::launch_step(P, s)); launches++;
    if (b->timing) CK(cudaEventRecord(next_event(b), s));
    if (sorted) { CK(ag::launch_order(b->d_cost, b->d_perm, b->N, P.sched, s)); b->perm_valid = true; launches++; }
    if (!P.obs_finish) { int rc = render_frame(b, 0, s, 1, fused); if (rc) return rc; launches++; }
    if (b->timing) CK(cudaEventRecord(next_event(b), s));
  } else {
    // the last num_frames ticks of the step each contribute one frame (the documented intent of
    // GridEnvironment::_partial_observation, GridEnvironment.hpp:413-433)
    if (b->frames > tps) CK(cudaMemsetAsync(b->d_obs, 0, b->obs_bytes, s));
    for (int t = 0; t < tps; t++) {
      P.n_ticks = 1; P.do_begin = (t == 0); P.do_end = 0;
      CK(ag::launch_step(P, s)); launches++;
      int frame = t - (tps - b->frames);
      if (frame >= 0) { int rc = render_frame(b, frame, s, 1); if (rc) return rc; launches++; }
    }
    P.n_ticks = 0; P.do_begin = 0; P.do_end = 1;
    CK(ag::launch_step(P, s)); launches++;
  }
  if (b->d_ram) {
    int rc = render_ram(b, s, 1); if (rc) return rc; launches++;
    if (b->timing && b->cfg.ram_obs == 2) CK(cudaEventRecord(next_event(b), s));  // (k_ram is this mode's observation kernel)
  }
  b->launches_last_step = launches;
  return refill_replay(b, s);
}

extern "C" int agarcl_batch_step(agarcl_batch* b, void* stream) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  return step_impl(b, (cudaStream_t)stream, false, nullptr);
}

extern "C" int agarcl_batch_set_timing(agarcl_batch* b, int enable) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  collect_timing(b);
  b->timing = enable != 0;
  b->acc_sim_ms = b->acc_obs_ms = 0.0;
  b->acc_steps = 0;
  return AGARCL_OK;
}
extern "C" int agarcl_batch_get_timing(agarcl_batch* b, double* sim_ms, double* obs_ms, int32_t* steps) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  collect_timing(b);
  if (sim_ms) *sim_ms = b->acc_sim_ms;
  if (obs_ms) *obs_ms = b->acc_obs_ms;
  if (steps) *steps = b->acc_steps;
  b->acc_sim_ms = b->acc_obs_ms = 0.0;
  b->acc_steps = 0;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_launches_per_step(const agarcl_batch* b) { return b ? b->launches_last_step : 0; }

extern "C" int agarcl_batch_costs(agarcl_batch* b, void* stream, uint32_t* out) {
  if (!b || !out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (!b->d_cost) return agarcl_set_error(AGARCL_ERR_STATE, "this batch keeps no per-instance costs");
  CK(cudaSetDevice(b->cfg.device));
  CK(cudaMemcpyAsync(out, b->d_cost, sizeof(uint32_t) * (size_t)b->N, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return AGARCL_OK;
}

extern "C" int agarcl_selftest_std_sort(const float* ys, int32_t n, uint16_t* order_out) {
  if (n < 0 || n > 65535 || (n > 0 && (!ys || !order_out))) return agarcl_set_error(AGARCL_ERR_INVALID, "bad keys / count");
  CK(ag::selftest_std_sort(ys, n, order_out));
  return AGARCL_OK;
}

extern "C" int agarcl_batch_flags(agarcl_batch* b, void* stream, uint32_t* or_all, uint32_t counts[32]) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  CK(cudaSetDevice(b->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  CK(ag::launch_flags(b->d_state, b->L, b->N, b->d_flagbuf, s));
  uint32_t h[33];
  CK(cudaMemcpyAsync(h, b->d_flagbuf, sizeof(h), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (or_all) *or_all = h[32];
  if (counts) std::memcpy(counts, h, 32 * sizeof(uint32_t));
  return AGARCL_OK;
}

extern "C" int agarcl_batch_obs(agarcl_batch* b, void** dev_ptr, int64_t shape[4], int32_t* dtype) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  if (dev_ptr) *dev_ptr = b->d_obs;
  if (shape) { shape[0] = (int64_t)b->N * b->A; shape[1] = (int64_t)b->frames * b->C; shape[2] = b->G; shape[3] = b->G; }
  if (dtype) *dtype = b->cfg.obs_dtype;
  return AGARCL_OK;
}
extern "C" int agarcl_batch_rewards(agarcl_batch* b, double** dev_ptr) {
  if (!b || !dev_ptr) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  *dev_ptr = b->d_rewards;
  return AGARCL_OK;
}
extern "C" int agarcl_batch_dones(agarcl_batch* b, uint8_t** dev_ptr) {
  if (!b || !dev_ptr) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  *dev_ptr = b->d_dones;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_step_host(agarcl_batch* b, const float* dxdy, const int32_t* act, void* obs_out,
                                      double* rewards_out, uint8_t* dones_out) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  int rc = agarcl_batch_set_actions(b, dxdy, act, 0, nullptr);
  if (rc) return rc;
  rc = agarcl_batch_step(b, nullptr);
  if (rc) return rc;
  const size_t NA = (size_t)b->N * b->A;
  if (obs_out) CK(cudaMemcpyAsync(obs_out, b->d_obs, b->obs_bytes, cudaMemcpyDeviceToHost, nullptr));
  if (rewards_out) CK(cudaMemcpyAsync(rewards_out, b->d_rewards, NA * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
  if (dones_out) CK(cudaMemcpyAsync(dones_out, b->d_dones, NA, cudaMemcpyDeviceToHost, nullptr));
  CK(cudaStreamSynchronize(nullptr));
  return AGARCL_OK;
}

// ---- host-resident observation mirror (mirror.cu)
static int ensure_mirror(agarcl_batch* b) {
  if (b->mirror) return AGARCL_OK;
  CK(cudaSetDevice(b->cfg.device));
  b->mirror = ag::mirror_create(b->N * b->A, b->A, b->frames * b->C, b->C, b->G, b->cfg.obs_dtype);
  return b->mirror ? AGARCL_OK : AGARCL_ERR_NOMEM;
}

extern "C" int agarcl_batch_sync_mirror(agarcl_batch* b, void* stream) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  int rc = ensure_mirror(b);
  if (rc) return rc;
  CK(cudaSetDevice(b->cfg.device));
  return ag::mirror_sync(b->mirror, b->d_obs, (cudaStream_t)stream);
}

extern "C" int agarcl_batch_mirror(agarcl_batch* b, void** host_ptr, int64_t shape[4], int32_t* dtype) {
  if (!b || !host_ptr) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  const bool fresh = b->mirror == nullptr;
  int rc = ensure_mirror(b);
  if (rc) return rc;
  if (fresh) {
    rc = ag::mirror_sync(b->mirror, b->d_obs, nullptr);
    if (rc) return rc;
  }
  *host_ptr = ag::mirror_ptr(b->mirror);
  if (shape) { shape[0] = (int64_t)b->N * b->A; shape[1] = (int64_t)b->frames * b->C; shape[2] = b->G; shape[3] = b->G; }
  if (dtype) *dtype = b->cfg.obs_dtype;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_step_mirror(agarcl_batch* b, const float* dxdy, const int32_t* act, double* rewards_out,
                                        uint8_t* dones_out) {
  if (!b) return agarcl_set_error(AGARCL_ERR_INVALID, "null batch");
  int rc = ensure_mirror(b);
  if (rc) return rc;
  if (!dxdy || !act) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  // the kernel reads the actions straight from host-mapped staging: no H2D copy in front of the launch
  const auto t0 = std::chrono::steady_clock::now();
  ag::mirror_stage_actions(b->mirror, dxdy, act, &b->cur_dxdy, &b->cur_act);
  bool lists_made = false;
  rc = step_impl(b, nullptr, true, &lists_made);
  if (rc) return rc;
  b->mirror_launch_us = (uint64_t)(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6);
  if (lists_made) {  // k_step lists what it scatters (and reward / done) into pinned host memory while it runs
    rc = ag::mirror_collect(b->mirror, b->d_obs, nullptr, rewards_out, dones_out);
    b->mirror_call_us = (uint64_t)(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6);
    return rc;
  }
  const size_t NA = (size_t)b->N * b->A;
  if (rewards_out) CK(cudaMemcpyAsync(rewards_out, b->d_rewards, NA * sizeof(double), cudaMemcpyDeviceToHost, nullptr));
  if (dones_out) CK(cudaMemcpyAsync(dones_out, b->d_dones, NA, cudaMemcpyDeviceToHost, nullptr));
  rc = ag::mirror_sync(b->mirror, b->d_obs, nullptr);
  b->launches_last_step += 1;  // k_pack
  b->mirror_call_us = (uint64_t)(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6);
  return rc;
}

extern "C" int agarcl_batch_step_lists(agarcl_batch* b, const float* dxdy, const int32_t* act, double* rewards_out, uint8_t* dones_out,
                                       agarcl_obs_lists* out) {
  if (!b || !dxdy || !act || !out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (b->frames != 1 || b->cfg.strict_reference || b->cfg.ram_obs == 2)
    return agarcl_set_error(AGARCL_ERR_STATE, "observation lists need the single fused step kernel (one frame, strict_reference = 0, a grid observation)");
  CK(cudaSetDevice(b->cfg.device));
  if (!b->lists) {
    b->lists = ag::mirror_create(b->N * b->A, b->A, b->frames * b->C, b->C, b->G, b->cfg.obs_dtype, true);
    if (!b->lists) return AGARCL_ERR_NOMEM;
  }
  ag::mirror_stage_actions(b->lists, dxdy, act, &b->cur_dxdy, &b->cur_act);
  bool lists_made = false;
  int rc = step_impl(b, nullptr, true, &lists_made, b->lists);
  if (rc) return rc;
  if (!lists_made) {
    CK(cudaStreamSynchronize(nullptr));
    return agarcl_set_error(AGARCL_ERR_STATE, "this observation configuration is not finished by the step kernel: no lists");
  }
  rc = ag::mirror_collect_lists(b->lists, nullptr, rewards_out, dones_out);
  if (rc) return rc;
  ag::mirror_lists_view(b->lists, out);
  return AGARCL_OK;
}

extern "C" int agarcl_batch_lists_expand(agarcl_batch* b, int32_t image, void* dense_out) {
  if (!b || !dense_out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (!b->lists) return agarcl_set_error(AGARCL_ERR_STATE, "no lists yet (agarcl_batch_step_lists)");
  if (image < 0 || image >= b->N * b->A) return agarcl_set_error(AGARCL_ERR_INVALID, "image out of range");
  if (ag::mirror_lists_expand(b->lists, image, dense_out) == 1) {  // overflowed its slot: the dense frame is on the device
    CK(cudaSetDevice(b->cfg.device));
    const size_t bytes = b->obs_bytes / ((size_t)b->N * b->A);
    CK(cudaMemcpy(dense_out, (const uint8_t*)b->d_obs + bytes * (size_t)image, bytes, cudaMemcpyDeviceToHost));
  }
  return AGARCL_OK;
}

extern "C" int agarcl_batch_mirror_stats(const agarcl_batch* b, uint64_t out[4]) {
  if (!b || !out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (!b->mirror) return agarcl_set_error(AGARCL_ERR_STATE, "no host mirror yet (agarcl_batch_mirror)");
  ag::MirrorStats st;
  ag::mirror_stats(b->mirror, &st);
  out[0] = st.entries; out[1] = st.dense_images; out[2] = st.d2h_bytes; out[3] = st.host_threads;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_mirror_timing(const agarcl_batch* b, uint64_t out[4]) {
  if (!b || !out) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (!b->mirror) return agarcl_set_error(AGARCL_ERR_STATE, "no host mirror yet (agarcl_batch_mirror)");
  ag::MirrorStats st;
  ag::mirror_stats(b->mirror, &st);
  out[0] = st.wait_us; out[1] = st.total_us; out[2] = b->mirror_launch_us; out[3] = b->mirror_call_us;
  return AGARCL_OK;
}

extern "C" int agarcl_batch_save_env_state(agarcl_batch* b, int32_t instance, const char* path) {
  if (!b || !path) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (instance < 0 || instance >= b->N) return agarcl_set_error(AGARCL_ERR_INVALID, "instance out of range");
  CK(cudaSetDevice(b->cfg.device));
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> blob(b->L.stride);
  CK(cudaMemcpy(blob.data(), b->d_state + (size_t)instance * b->L.stride, b->L.stride, cudaMemcpyDeviceToHost));
  return agarcl_snapshot_write(&b->cfg, &b->L, blob.data(), path);
}

extern "C" int agarcl_batch_load_env_state(agarcl_batch* b, int32_t instance, const char* path, int lossless) {
  if (!b || !path) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (instance < 0 || instance >= b->N) return agarcl_set_error(AGARCL_ERR_INVALID, "instance out of range");
  if (!b->was_reset) return agarcl_set_error(AGARCL_ERR_STATE, "load_env_state() before reset()");
  CK(cudaSetDevice(b->cfg.device));
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> blob(b->L.stride);
  CK(cudaMemcpy(blob.data(), b->d_state + (size_t)instance * b->L.stride, b->L.stride, cudaMemcpyDeviceToHost));
  int rc = agarcl_snapshot_read(&b->cfg, &b->L, blob.data(), path, lossless);
  if (rc) return rc;
  const agarcl_inst_hdr* hdr = reinterpret_cast<const agarcl_inst_hdr*>(blob.data() + b->L.off_hdr);
  b->seeds[instance] = (uint64_t)hdr->seed_lo | ((uint64_t)hdr->seed_hi << 32);
  CK(cudaMemcpy(b->d_seeds + instance, &b->seeds[instance], sizeof(uint64_t), cudaMemcpyHostToDevice));
  if (b->cfg.rng_mode == AGARCL_RNG_MT19937) {  // Engine::seed(agarcl_data["seed"]): the mt19937_64 stream restarts (at the saved cursor when lossless)
    rc = mt_restart(b, instance, hdr->seed_lo, hdr->rng_cursor);
    if (rc) return rc;
  }
  CK(cudaMemcpy(b->d_state + (size_t)instance * b->L.stride, blob.data(), b->L.stride, cudaMemcpyHostToDevice));
  return AGARCL_OK;
}

extern "C" int agarcl_batch_download_state(agarcl_batch* b, int32_t instance, void* blob) {
  if (!b || !blob) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (instance < 0 || instance >= b->N) return agarcl_set_error(AGARCL_ERR_INVALID, "instance out of range");
  CK(cudaSetDevice(b->cfg.device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(blob, b->d_state + (size_t)instance * b->L.stride, b->L.stride, cudaMemcpyDeviceToHost));
  return AGARCL_OK;
}
extern "C" int agarcl_batch_upload_state(agarcl_batch* b, int32_t instance, const void* blob) {
  if (!b || !blob) return agarcl_set_error(AGARCL_ERR_INVALID, "null argument");
  if (instance < 0 || instance >= b->N) return agarcl_set_error(AGARCL_ERR_INVALID, "instance out of range");
  CK(cudaSetDevice(b->cfg.device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(b->d_state + (size_t)instance * b->L.stride, blob, b->L.stride, cudaMemcpyHostToDevice));
  b->was_reset = true;
  return AGARCL_OK;
}
