// layout.cpp — host-side derivation of the per-instance state layout and of everything that is a
// pure function of the environment configuration (mode flags, bot roster, player iteration order).
//
// Reference behaviour restated here:
//   Engine::set_mode                     agario/engine/Engine.hpp:367-416
//   BaseEnvironment::reset / add_bots    environment/envs/BaseEnvironment.hpp:179-204,374-425
//   create_squared_pellets (count only)  agario/engine/Engine.hpp:426-475
//   GridObservation::channels_per_frame  environment/envs/GridEnvironment.hpp:188-195
//   GameState::PlayerMap iteration       agario/engine/GameState.hpp:44  (std::unordered_map<pid, ...>)
#include <cstring>
#include <random>
#include <unordered_map>

#include "../../include/agarcl_b200.h"
#include "host_util.h"

static uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

extern "C" int agarcl_make_layout(const agarcl_cfg* c, agarcl_layout* L) {
  if (!c || !L) return agarcl_set_error(AGARCL_ERR_INVALID, "null cfg/layout");
  std::memset(L, 0, sizeof(*L));
  if (c->num_agents < 1) return agarcl_set_error(AGARCL_ERR_INVALID, "num_agents must be >= 1");
  if (c->ticks_per_step < 1) return agarcl_set_error(AGARCL_ERR_INVALID, "ticks_per_step must be a positive integer");
  if (c->arena_size < 8 || c->arena_size > 16384) return agarcl_set_error(AGARCL_ERR_INVALID, "arena_size out of range [8,16384]");
  if (c->num_pellets < 0 || c->num_pellets > 65000) return agarcl_set_error(AGARCL_ERR_INVALID, "num_pellets out of range [0,65000]");
  if (c->num_viruses < 0 || c->num_viruses > 4096) return agarcl_set_error(AGARCL_ERR_INVALID, "num_viruses out of range");
  if (c->num_bots < 0) return agarcl_set_error(AGARCL_ERR_INVALID, "num_bots must be >= 0");
  if (c->mode_number < 0 || c->mode_number > 10) return agarcl_set_error(AGARCL_ERR_INVALID, "Invalid mode number");
  if (c->grid_size < 1 || c->grid_size > 1024) return agarcl_set_error(AGARCL_ERR_INVALID, "grid_size out of range");
  if (c->num_frames < 1) return agarcl_set_error(AGARCL_ERR_INVALID, "num_frames must be >= 1");

  const int mode = c->mode_number;
  // Engine::set_mode
  int base = mode;
  if (mode == 5) base = 2;
  if (mode == 6 || mode >= 7) base = 4;
  L->mass_decay = (base == 0 || base == 2 || base == 4);
  L->squared_pellets = (base == 1 || base == 2);
  L->regen = (base == 0 || base == 3 || base == 4);
  L->agent_mass = (mode == 5 || mode == 6) ? 1000 : 25;

  // BaseEnvironment::reset: bots only in mode 0 (num_bots of them) and modes 7..10 (exactly one)
  int bots = 0;
  if (mode == 0) bots = c->num_bots;
  else if (mode > 6) bots = 1;
  L->A = c->num_agents;
  L->P = c->num_agents + bots;
  if (L->P > AGARCL_MAX_PLAYERS) return agarcl_set_error(AGARCL_ERR_INVALID, "too many players per instance (agents + bots > 64)");
  for (int p = 0; p < L->P; p++) {
    if (p < L->A) L->bot_type[p] = -1;
    else if (mode == 0) {
      int i = p - L->A;  // add_bots: switch (i % num_bots_) with i < num_bots_  (quirk Q14)
      L->bot_type[p] = (i < 4) ? i : 0;
    } else {
      int idx = mode - 7;  // custom_add_bot
      L->bot_type[p] = (idx >= 0 && idx < 4) ? idx : 0;
    }
  }
  // player iteration order: the same container the reference iterates, fed the same insert sequence
  {
    std::unordered_map<unsigned short, int> m;
    for (int p = 0; p < L->P; p++) m.insert(std::make_pair((unsigned short)p, p));
    int k = 0;
    for (auto& kv : m) L->order[k++] = kv.second;
  }

  L->cap_cells = AGARCL_MAX_CELLS;
  int cap_p = c->num_pellets;
  if (L->squared_pellets) {
    int sq = 4 * (c->arena_size / 2);  // 4 sides x int(min(W,H)/2 / spacing 1)
    if (sq > cap_p) cap_p = sq;
  }
  if (cap_p < 1) cap_p = 1;
  if (cap_p > 65000) return agarcl_set_error(AGARCL_ERR_INVALID, "pellet capacity exceeds 65000");
  L->cap_pellets = cap_p;
  L->cap_viruses = c->cap_viruses > 0 ? c->cap_viruses : c->num_viruses + 32;
  if (L->cap_viruses < c->num_viruses) return agarcl_set_error(AGARCL_ERR_INVALID, "cap_viruses < num_viruses");
  L->cap_foods = c->cap_foods > 0 ? c->cap_foods : 256;
  L->cap_replay = c->cap_replay > 0 ? c->cap_replay : ((c->rng_mode == AGARCL_RNG_PHILOX) ? 0 : 16384);
  L->obs_channels = 1 + (c->observe_cells != 0) + 2 * (c->observe_others != 0) + 2 * (c->observe_viruses != 0) +
                    2 * (c->observe_pellets != 0);

  uint32_t off = 0;
  L->off_hdr = off;      off += (uint32_t)sizeof(agarcl_inst_hdr);
  L->off_players = off;  off += (uint32_t)sizeof(agarcl_player) * (uint32_t)L->P;
  L->off_cells = off;    off += (uint32_t)sizeof(agarcl_cell) * (uint32_t)L->P * (uint32_t)L->cap_cells;
  L->off_viruses = off;  off += (uint32_t)sizeof(agarcl_virus) * (uint32_t)L->cap_viruses;
  L->off_foods = off;    off += (uint32_t)sizeof(agarcl_food) * (uint32_t)L->cap_foods;
  off = align_up(off, 16);
  L->off_pellets = off;  off += (uint32_t)sizeof(agarcl_pellet) * (uint32_t)L->cap_pellets;
  L->stride = align_up(off, 128);
  return AGARCL_OK;
}

// std::mt19937_64 + std::uniform_real_distribution<float>: the generator the reference seeds in
// Engine::seed (Engine.hpp:242-245) and draws from in random_location (Engine.hpp:143-148,1304-1311).
extern "C" int agarcl_mt19937_draws(uint64_t seed, float* out, int32_t n) {
  if (!out || n < 0) return agarcl_set_error(AGARCL_ERR_INVALID, "bad draws buffer");
  std::mt19937_64 rng((unsigned)seed);  // BaseEnvironment::seed(int) -> Engine::seed(unsigned)
  for (int i = 0; i < n; i++) {
    std::uniform_real_distribution<float> d(0.0f, 1.0f);
    out[i] = d(rng);
  }
  return AGARCL_OK;
}
