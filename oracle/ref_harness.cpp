/*
 * ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C-ABI wrapper around the UNMODIFIED reference engine compiled from the sources where
 * they lie under /root/reference (see oracle/Makefile).  It drives the reference's own
 * agario::env::GridEnvironment<int,false> (environment/envs/GridEnvironment.hpp:350-491)
 * and dumps its state into the agarcl_b200 state blob so tests can compare the oracle port
 * (oracle/oracle.c) and the CUDA path against the real thing.
 *
 * Declared deviations from the shipped reference (DESIGN.md "Oracle"):
 *  1. GridEnvironment.hpp:374 initialises an OpenGL member that only exists under RENDERABLE;
 *     the Makefile drops that one mem-initialiser in a scratch copy (no behavioural effect).
 *  2. The recombine timer reads std::chrono::steady_clock (Entities.hpp:127,185,190), i.e. WALL
 *     time.  Here `steady_clock` is re-pointed (preprocessor, no source edit) at a simulation
 *     clock that counts engine ticks at 30 ticks/s, so RECOMBINE_TIMER_SEC = 10 s = 300 ticks.
 *  3. Each reset starts from a fresh player map and pid 0 ("fresh Engine per episode").
 */
#include <algorithm>
#include <any>
#include <array>
#include <bitset>
#include <cctype>
#include <cerrno>
#include <cinttypes>
#include <ciso646>
#include <clocale>
#include <cstddef>
#include <cstdio>
#include <deque>
#include <exception>
#include <filesystem>
#include <forward_list>
#include <initializer_list>
#include <iosfwd>
#include <istream>
#include <iterator>
#include <list>
#include <map>
#include <optional>
#include <ostream>
#include <queue>
#include <string_view>
#include <type_traits>
#include <typeinfo>
#include <utility>
#include <valarray>
#include <variant>
#include <version>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <atomic>

#include "../include/agarcl_b200.h"

/* ---- deviation 2: simulation clock standing in for steady_clock ------------------------- */
namespace std { namespace chrono {
struct agarcl_sim_clock {
  typedef std::chrono::duration<long long, std::ratio<1, 30>> duration; /* one engine tick */
  typedef duration::rep rep;
  typedef duration::period period;
  typedef std::chrono::time_point<agarcl_sim_clock, duration> time_point;
  static constexpr bool is_steady = true;
  static thread_local const unsigned long* tick_ptr;
  static time_point now() noexcept { return time_point(duration(tick_ptr ? (long long)*tick_ptr : 0LL)); }
};
thread_local const unsigned long* agarcl_sim_clock::tick_ptr = nullptr;
}}

typedef int screen_len;
class FBOException : public std::runtime_error { using runtime_error::runtime_error; };

#define steady_clock agarcl_sim_clock
#define private public
#define protected public
#include <environment/envs/GridEnvironment.hpp> /* scratch copy in oracle/_ref/gen shadows line 374 */
/* GoBiggerObservation only: the scratch copy in oracle/_ref/gen stops before the GoBiggerEnvironment class
 * (it needs the OpenGL frame buffer); the header's `#define f first` / `#define s second` are undone. */
#include <environment/envs/GoBiggerEnvironment.hpp>
#undef f
#undef s
#undef private
#undef protected
#undef steady_clock

using RefEnvT = agario::env::GridEnvironment<int, false>;
using RefObsT = agario::env::GridObservation<int, false>;
using RefRamT = agario::env::GoBiggerObservation<false>;
using SimClock = std::chrono::agarcl_sim_clock;

struct RefEnv {
  std::unique_ptr<RefEnvT> env;
  std::unique_ptr<RefObsT> forced; /* add_frame(...,0) target, see Q11 */
  std::unique_ptr<RefRamT> ram;    /* GoBiggerObservation fed with this engine's state */
  int arena;
  int num_agents, grid, channels;
  /* pid of player index 0 in the current episode: 0 behind ref_reset (deviation 3), e * P behind ref_reset_native (Q3) */
  int pid_base() const {
    int lo = 1 << 30;
    for (auto& pr : env->engine_.state.players) lo = std::min(lo, (int)pr.first);
    return lo == (1 << 30) ? 0 : lo;
  }
  void bind_clock() { SimClock::tick_ptr = &env->engine_.state.ticks; }
};

struct CoutSilencer {
  std::streambuf* old;
  std::ostringstream sink;
  CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~CoutSilencer() { std::cout.rdbuf(old); }
};

static void fresh_reset(RefEnv* r) {
  /* deviation 3: fresh player map + pid 0 so that map iteration order is that of a new Engine */
  auto& st = r->env->engine_.state;
  st.players = decltype(st.players)();
  st.next_pid = 0;
  r->env->reset();
}

static bool g_pool_silenced = false;  /* ref_pool_run has redirected cout / cerr for all its worker threads */
struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };

extern "C" {

void* ref_create(const agarcl_cfg* c) {
  auto* r = new RefEnv();
  {
    CoutSilencer quiet;
    SimClock::tick_ptr = nullptr;
    r->env.reset(new RefEnvT(c->num_agents, c->ticks_per_step, c->arena_size, c->pellet_regen != 0, c->num_pellets,
                             c->num_viruses, c->num_bots, c->reward_type, c->c_death, c->mode_number));
  }
  r->bind_clock();
  r->env->configure_observation(c->num_frames, c->grid_size, c->observe_cells != 0, c->observe_others != 0,
                                c->observe_viruses != 0, c->observe_pellets != 0);
  r->forced.reset(new RefObsT(1, c->grid_size, c->observe_cells != 0, c->observe_others != 0,
                              c->observe_viruses != 0, c->observe_pellets != 0));
  r->num_agents = c->num_agents;
  r->arena = c->arena_size;
  r->grid = c->grid_size;
  r->channels = std::get<0>(r->forced->shape());
  return r;
}

void ref_destroy(void* h) {
  SimClock::tick_ptr = nullptr;
  delete static_cast<RefEnv*>(h);
}

void ref_seed(void* h, unsigned s) { static_cast<RefEnv*>(h)->env->seed((int)s); }

void ref_reset(void* h) {
  auto* r = static_cast<RefEnv*>(h);
  r->bind_clock();
  fresh_reset(r);
}

/* BaseEnvironment::reset exactly as the reference runs it: the player map is cleared, not replaced, and next_pid keeps counting
 * (quirk Q3) -- the episode's pids are e * P .. e * P + P - 1 and the map's iteration order is that of the reused bucket array */
void ref_reset_native(void* h) {
  auto* r = static_cast<RefEnv*>(h);
  r->bind_clock();
  CoutSilencer quiet;
  r->env->reset();
}
int ref_pid_base(void* h) { return static_cast<RefEnv*>(h)->pid_base(); }

void ref_take_actions(void* h, const float* dxdy, const int32_t* act) {
  auto* r = static_cast<RefEnv*>(h);
  r->bind_clock();
  std::vector<agario::env::Action> a;
  for (int i = 0; i < r->num_agents; i++)
    a.emplace_back(dxdy[2 * i], dxdy[2 * i + 1], static_cast<agario::action>(act[i]));
  r->env->take_actions(a);
}

/* rewards in the reference's own return order (map order of non-bot players, quirk Q15) */
void ref_step(void* h, double* rewards) {
  auto* r = static_cast<RefEnv*>(h);
  r->bind_clock();
  auto rw = r->env->step();
  for (size_t i = 0; i < rw.size(); i++) rewards[i] = rw[i];
}

void ref_dones(void* h, uint8_t* out) {
  auto* r = static_cast<RefEnv*>(h);
  auto d = r->env->dones();
  for (size_t i = 0; i < d.size(); i++) out[i] = d[i] ? 1 : 0;
}

/* GridObservation::add_frame(player, state, 0) on a cleared buffer (GridEnvironment.hpp:91-123) */
void ref_obs(void* h, int agent, int32_t* out) {
  auto* r = static_cast<RefEnv*>(h);
  r->bind_clock();
  auto& player = r->env->engine_.player(r->env->pids_[agent]);
  r->forced->clear_data();
  r->forced->add_frame(player, r->env->engine_.game_state(), 0);
  std::memcpy(out, r->forced->data(), sizeof(int32_t) * (size_t)r->forced->length());
}

/* GoBiggerObservation::add_frame(player 0, state, 0) (GoBiggerEnvironment.hpp:515-548), flattened into the
 * agarcl ram records [P][AGARCL_RAM_RECORD].  The observation object persists between calls, so players
 * with nothing in view keep their previous PlayerState exactly as in the reference; ref_ram_clear is
 * GoBiggerEnvironment::reset's observation.clear(). */
void ref_ram_clear(void* h) {
  auto* r = static_cast<RefEnv*>(h);
  if (r->ram) r->ram->clear();
}
void ref_ram_obs(void* h, int P, float* out) {
  auto* r = static_cast<RefEnv*>(h);
  r->bind_clock();
  /* GoBiggerObservation chats on cout / cerr ("Player id not found ..."): silenced per call, or once around a whole
   * multi-threaded pool run (the stream buffers are process-wide: swapping them from several threads is a race) */
  std::unique_ptr<CoutSilencer> quiet;
  std::streambuf* olderr = nullptr;
  if (!g_pool_silenced) { quiet.reset(new CoutSilencer()); olderr = std::cerr.rdbuf(quiet->sink.rdbuf()); }
  if (!r->ram) {
    r->ram.reset(new RefRamT(r->arena, r->arena, 0, 0, r->num_agents));
    r->ram->configure(1, r->grid, true, true, true, true);
  }
  auto& player = r->env->engine_.player(r->env->pids_[0]);
  r->ram->add_frame(player, r->env->engine_.game_state(), 0);
  if (olderr) std::cerr.rdbuf(olderr);
  const int pid_base = r->pid_base();
  for (auto& kv : r->ram->get_player_states().get_all_player_states()) {
    int pid = (int)kv.first - pid_base;
    if (pid < 0 || pid >= P) continue;
    const auto& ps = kv.second;
    float* rec = out + (size_t)pid * AGARCL_RAM_RECORD;
    std::memset(rec, 0, sizeof(float) * AGARCL_RAM_RECORD);
    const auto& fo = ps.get_food_infos();
    const auto& vi = ps.get_virus_infos();
    const auto& sp = ps.get_spore_infos();
    const auto& cl = ps.get_clone_infos();
    for (size_t i = 0; i < fo.size() && i < AGARCL_RAM_KP; i++) {
      float* e = rec + AGARCL_RAM_OFF_FOOD + 4 * i;
      e[0] = fo[i].position.x; e[1] = fo[i].position.y; e[2] = (float)fo[i].radius; e[3] = (float)fo[i].score;
    }
    for (size_t i = 0; i < vi.size() && i < AGARCL_RAM_KV; i++) {
      float* e = rec + AGARCL_RAM_OFF_VIRUS + 4 * i;
      e[0] = vi[i].position.x; e[1] = vi[i].position.y; e[2] = (float)vi[i].radius; e[3] = (float)vi[i].score;
    }
    for (size_t i = 0; i < sp.size() && i < AGARCL_RAM_KS; i++) {
      float* e = rec + AGARCL_RAM_OFF_SPORE + 4 * i;
      e[0] = sp[i].position.x; e[1] = sp[i].position.y; e[2] = (float)sp[i].radius; e[3] = (float)sp[i].score;
    }
    for (size_t i = 0; i < cl.size() && i < AGARCL_RAM_KC; i++) {
      float* e = rec + AGARCL_RAM_OFF_CLONE + 8 * i;
      e[0] = cl[i].position.x; e[1] = cl[i].position.y; e[2] = (float)cl[i].radius; e[3] = (float)cl[i].score;
      e[4] = (float)cl[i].velocity.first; e[5] = (float)cl[i].velocity.second; e[6] = cl[i].direction; e[7] = (float)cl[i].owner;
    }
    rec[0] = (float)fo.size(); rec[1] = (float)vi.size(); rec[2] = (float)sp.size(); rec[3] = (float)cl.size();
    rec[4] = (float)ps.get_score();
    rec[7] = (float)((fo.size() > AGARCL_RAM_KP ? 1 : 0) | (vi.size() > AGARCL_RAM_KV ? 2 : 0) | (sp.size() > AGARCL_RAM_KS ? 4 : 0) |
                     (cl.size() > AGARCL_RAM_KC ? 8 : 0));
  }
}

/* pids_ : which player each agent index drives.  After load_env_state the reference rebuilds it by iterating
 * the player map (BaseEnvironment.hpp:330-339), i.e. in REVERSED agent order for the usual small rosters. */
int ref_agent_pids(void* h, int32_t* out) {
  auto* r = static_cast<RefEnv*>(h);
  int k = 0;
  for (auto pid : r->env->pids_) out[k++] = (int32_t)pid;
  return k;
}

/* BaseEnvironment::save_env_state / load_env_state (BaseEnvironment.hpp:213-343) on the reference itself */
int ref_save_env_state(void* h, const char* path) {
  auto* r = static_cast<RefEnv*>(h);
  try { r->env->save_env_state(path); } catch (const std::exception&) { return -1; }
  return 0;
}
int ref_load_env_state(void* h, const char* path) {
  auto* r = static_cast<RefEnv*>(h);
  r->bind_clock();
  CoutSilencer quiet;
  try { r->env->load_env_state(path); } catch (const std::exception&) { return -1; }
  r->env->is_loading_env_state = false;  /* the flag only suppresses the reset() of the constructor */
  return 0;
}

/* what the reference's own step()/get_state() left in its buffer (quirk Q11) */
void ref_obs_native(void* h, int agent, int32_t* out) {
  auto* r = static_cast<RefEnv*>(h);
  auto& obs = r->env->get_observations()[agent];
  std::memcpy(out, obs.data(), sizeof(int32_t) * (size_t)obs.length());
}
int ref_obs_native_len(void* h) { return static_cast<RefEnv*>(h)->env->get_observations()[0].length(); }

/* map iteration order of players (GameState.hpp:44) as pids */
int ref_player_order(void* h, int32_t* out) {
  auto* r = static_cast<RefEnv*>(h);
  int k = 0;
  const int base = r->pid_base();
  for (auto& pr : r->env->engine_.state.players) out[k++] = (int)pr.first - base;
  return k;
}

/* next n canonical draws the engine's rng will produce, from a COPY of it (GameState.hpp:51;
 * uniform_real_distribution<float>(0,1) consumes exactly one 64-bit word per draw) */
void ref_rng_peek(void* h, float* out, int n) {
  auto* r = static_cast<RefEnv*>(h);
  std::mt19937_64 copy = r->env->engine_.state.rng;
  for (int i = 0; i < n; i++) {
    std::uniform_real_distribution<float> d(0.0f, 1.0f);
    out[i] = d(copy);
  }
}

void ref_set_cell_mass(void* h, int pid, int cell, unsigned mass) {
  auto* r = static_cast<RefEnv*>(h);
  r->env->engine_.player((agario::pid)(pid + r->pid_base())).cells.at(cell).set_mass(mass);  /* pid = player index */
}
void ref_set_cell_pos(void* h, int pid, int cell, float x, float y) {
  auto* r = static_cast<RefEnv*>(h);
  auto& c = r->env->engine_.player((agario::pid)(pid + r->pid_base())).cells.at(cell);
  c.x = x;
  c.y = y;
}
void ref_set_virus(void* h, int idx, float x, float y) {
  auto* r = static_cast<RefEnv*>(h);
  auto& v = r->env->engine_.state.viruses.at(idx);
  v.x = x;
  v.y = y;
}

/* Dump the whole GameState into one agarcl blob.  Returns 0, or a negative count of capacity misses. */
int ref_dump_state(void* h, const agarcl_layout* L, void* blob_) {
  auto* r = static_cast<RefEnv*>(h);
  auto& eng = r->env->engine_;
  auto& st = eng.state;
  auto* blob = static_cast<uint8_t*>(blob_);
  std::memset(blob, 0, L->stride);
  int miss = 0;
  auto* hdr = reinterpret_cast<agarcl_inst_hdr*>(blob + L->off_hdr);
  hdr->tick = (uint32_t)st.ticks;
  hdr->n_pellets = (int32_t)st.pellets.size();
  hdr->n_viruses = (int32_t)st.viruses.size();
  hdr->n_foods = (int32_t)st.foods.size();
  hdr->done_sticky = r->env->dones_[0] ? 1u : 0u;
  auto* pel = reinterpret_cast<agarcl_pellet*>(blob + L->off_pellets);
  for (size_t i = 0; i < st.pellets.size(); i++) {
    if ((int)i >= L->cap_pellets) { miss--; break; }
    pel[i].x = st.pellets[i].x;
    pel[i].y = st.pellets[i].y;
  }
  auto* vir = reinterpret_cast<agarcl_virus*>(blob + L->off_viruses);
  for (size_t i = 0; i < st.viruses.size(); i++) {
    if ((int)i >= L->cap_viruses) { miss--; break; }
    vir[i].x = st.viruses[i].x;
    vir[i].y = st.viruses[i].y;
    vir[i].vx = st.viruses[i].velocity.dx;
    vir[i].vy = st.viruses[i].velocity.dy;
    vir[i].mass = st.viruses[i].mass();
    vir[i].hits = st.viruses[i].get_num_food_hits();
  }
  auto* foo = reinterpret_cast<agarcl_food*>(blob + L->off_foods);
  for (size_t i = 0; i < st.foods.size(); i++) {
    if ((int)i >= L->cap_foods) { miss--; break; }
    foo[i].x = st.foods[i].x;
    foo[i].y = st.foods[i].y;
    foo[i].vx = st.foods[i].velocity.dx;
    foo[i].vy = st.foods[i].velocity.dy;
  }
  auto* pls = reinterpret_cast<agarcl_player*>(blob + L->off_players);
  auto* cells = reinterpret_cast<agarcl_cell*>(blob + L->off_cells);
  uint32_t max_id = 0;
  const int pid_base = r->pid_base();
  for (auto& pr : st.players) {
    int p = (int)pr.first - pid_base;
    if (p >= L->P) { miss--; continue; }
    auto& pl = *pr.second;
    agarcl_player& o = pls[p];
    o.n_cells = (int32_t)pl.cells.size();
    o.target_x = pl.target.x;
    o.target_y = pl.target.y;
    o.action = (int32_t)pl.action;
    o.split_cd = (int32_t)pl.split_cooldown;
    o.feed_cd = (int32_t)pl.feed_cooldown;
    o.anti_team_decay = pl.anti_team_decay;
    o.elapsed_ticks = pl.elapsed_ticks;
    o.last_decay_tick = pl.last_decay_tick;
    o.bot_type = L->bot_type[p];
    o.min_mass_cell = pl._minMassCell;
    o.food_eaten = pl.food_eaten;
    o.highest_mass = pl.highest_mass;
    o.cells_eaten = pl.cells_eaten;
    o.viruses_eaten = pl.viruses_eaten;
    o.vet_count = (int32_t)pl.virus_eaten_ticks.size();
    for (size_t k = 0; k < pl.virus_eaten_ticks.size() && k < AGARCL_VET_CAP; k++) o.vet_ticks[k] = pl.virus_eaten_ticks[k];
    if (pl.virus_eaten_ticks.size() > AGARCL_VET_CAP) miss--;
    for (size_t k = 0; k < pl.cells.size(); k++) {
      if ((int)k >= L->cap_cells) { miss--; break; }
      auto& c = pl.cells[k];
      agarcl_cell& oc = cells[(size_t)p * L->cap_cells + k];
      oc.x = c.x;
      oc.y = c.y;
      oc.vx = c.velocity.dx;
      oc.vy = c.velocity.dy;
      oc.svx = c.splitting_velocity.dx;
      oc.svy = c.splitting_velocity.dy;
      oc.mass = c.mass();
      oc.id = (uint32_t)c.id;
      long long rt = c._recombine_timer.time_since_epoch().count();
      oc.recomb_tick = rt < 0 ? 0u : (uint32_t)rt;
      max_id = std::max(max_id, (uint32_t)c.id);
    }
  }
  hdr->next_cell_id = max_id + 1;
  return miss;
}

/* iteration order of a real std::unordered_map<int, ...> after inserting keys in order (pins oracle_umap_order) */
/* the real std::sort on the strips' element type with the comparator of collision_detection.hpp:29-31 (pins oracle.c's restatement) */
void ref_std_sort_pairs(int* ids, float* ys, int n) {
  std::vector<std::pair<int, float>> v;
  for (int i = 0; i < n; i++) v.emplace_back(std::make_pair(ids[i], ys[i]));
  std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
  for (int i = 0; i < n; i++) { ids[i] = v[i].first; ys[i] = v[i].second; }
}

void ref_umap_order(const int* keys, int n, int* out) {
  std::unordered_map<int, std::vector<int>> m;
  for (int i = 0; i < n; i++) m[keys[i]].push_back(i);
  int k = 0;
  for (auto& kv : m) out[k++] = kv.first;
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline: M independent GridEnvironment instances, one engine per worker thread with no
 * exchange — the pattern of BotEvaluator::run (agario/bots/benchmark.cpp:146-168): each work item
 * runs `steps` env-steps of (random actions, step(), forced add_frame) of one instance.
 * The workers are plain std::threads drawing instances from an atomic counter, NOT the reference's
 * utils/thread-pool.h: that pool starts its workers while the constructor is still growing the
 * vectors they index (thread-pool.cpp:18-26) and its destructor signals the dispatcher without the
 * mutex (:112-114, a lost wake-up) — it hung a whole bench run on a 16-thread box.
 * Environments are built once (ref_pool_create) so that only stepping is timed.
 * ------------------------------------------------------------------------------------------ */
struct RefPool {
  agarcl_cfg cfg;
  std::vector<RefEnv*> envs;
  std::vector<std::mt19937> arng;
  float p_feed = -1.0f, p_split = -1.0f;  /* < 0: a ~ U{0,1,2} (bench/go_bigger_example.py:100-103) */
};

void* ref_pool_create(const agarcl_cfg* c, int instances, unsigned base_seed) {
  auto* p = new RefPool();
  p->cfg = *c;
  for (int i = 0; i < instances; i++) {
    RefEnv* r = static_cast<RefEnv*>(ref_create(c));
    ref_seed(r, base_seed + (unsigned)i);
    ref_reset(r);
    p->envs.push_back(r);
    p->arng.emplace_back(base_seed * 7919u + (unsigned)i);
  }
  return p;
}
/* action mix of the pool's random-walk policy (feed w.p. p_feed, split w.p. p_split, else none) and, with boost > 0,
 * every agent's first cell raised to that mass with Cell::set_mass (BASELINE configs[3]) */
void ref_pool_profile(void* h, float p_feed, float p_split, unsigned boost) {
  auto* p = static_cast<RefPool*>(h);
  p->p_feed = p_feed;
  p->p_split = p_split;
  if (boost)
    for (auto* r : p->envs)
      for (int a = 0; a < p->cfg.num_agents; a++) ref_set_cell_mass(r, a, 0, boost);
}
void ref_pool_destroy(void* h) {
  auto* p = static_cast<RefPool*>(h);
  for (auto* r : p->envs) ref_destroy(r);
  delete p;
}
/* returns wall seconds for `steps` env-steps of every instance on `threads` pool workers */
double ref_pool_run(void* h, int threads, int steps, int with_obs) {
  auto* p = static_cast<RefPool*>(h);
  const agarcl_cfg* c = &p->cfg;
  auto work = [p, c, steps, with_obs](int i) {
    std::vector<float> dxdy((size_t)c->num_agents * 2);
    std::vector<int32_t> act((size_t)c->num_agents);
    std::vector<double> rew((size_t)c->num_agents);
    RefEnv* r = p->envs[i];
    std::vector<int32_t> obs((size_t)r->forced->length());
    std::mt19937& arng = p->arng[i];
    std::uniform_real_distribution<float> u(-1.0f, 1.0f), u01(0.0f, 1.0f);
    std::vector<float> ram;
    const int P = c->num_agents + c->num_bots;
    if (with_obs == 2) ram.resize((size_t)P * AGARCL_RAM_RECORD);
    for (int s = 0; s < steps; s++) {
      for (int a = 0; a < c->num_agents; a++) {
        dxdy[2 * a] = u(arng);
        dxdy[2 * a + 1] = u(arng);
        if (p->p_feed < 0.0f) act[a] = (int)(arng() % 3u);
        else { const float v = u01(arng); act[a] = v < p->p_feed ? 1 : (v < p->p_feed + p->p_split ? 2 : 0); }
      }
      ref_take_actions(r, dxdy.data(), act.data());
      ref_step(r, rew.data());
      if (with_obs == 1)
        for (int a = 0; a < c->num_agents; a++) ref_obs(r, a, obs.data());
      else if (with_obs == 2)  /* GoBiggerObservation::add_frame for every player + the records (the "ram" observation) */
        ref_ram_obs(r, P, ram.data());
    }
  };
  NullBuf nullbuf;  /* (a streambuf that drops everything is safe to share between the workers) */
  std::streambuf *oldout = nullptr, *olderr = nullptr;
  if (with_obs == 2) { oldout = std::cout.rdbuf(&nullbuf); olderr = std::cerr.rdbuf(&nullbuf); g_pool_silenced = true; }
  auto t0 = std::chrono::high_resolution_clock::now();
  {
    std::atomic<int> next(0);
    const int n = (int)p->envs.size();
    std::vector<std::thread> workers;
    for (int t = 0; t < threads; t++)
      workers.emplace_back([&work, &next, n]() {
        for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) work(i);
      });
    for (auto& w : workers) w.join();
  }
  auto t1 = std::chrono::high_resolution_clock::now();
  if (with_obs == 2) { std::cout.rdbuf(oldout); std::cerr.rdbuf(olderr); g_pool_silenced = false; }
  return std::chrono::duration<double>(t1 - t0).count();
}

double ref_bench(const agarcl_cfg* c, int instances, int threads, int steps, int with_obs, unsigned base_seed) {
  void* p = ref_pool_create(c, instances, base_seed);
  double sec = ref_pool_run(p, threads, steps, with_obs);
  ref_pool_destroy(p);
  return (double)instances * (double)steps / sec;
}

} /* extern "C" */
