// pybind_agarcl.cpp — the Python module `agarcl` of the reference (environment/bindings.cpp:94-135,181-374), rebuilt over the
// C ABI of libagarcl_b200.so: same module name, class names, constructor signatures, method names and return shapes, so that
// gym_agario/AgarioEnv.py (`import agarcl`) runs on the CUDA library unchanged.  A GridEnvironment / GoBiggerEnvironment is a
// size-1 batch; every call goes through include/agarcl_b200.h with HOST buffers (no CUDA headers here).  Nothing below computes
// game logic.  Screen environments need OpenGL and are out of scope: has_screen_env is False, like a reference build without it
// (bindings.cpp:172-176).  Built by agarcl_b200/build.py into agarcl_b200/agarcl.<abi>.so.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <sstream>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "../../include/agarcl_b200.h"

namespace py = pybind11;

namespace {

void ck(int rc) {  // EnvironmentException / EngineException of the reference surface as RuntimeError (BaseEnvironment.hpp:19)
  if (rc) throw std::runtime_error(agarcl_last_error());
}

// BaseEnvironment + GridEnvironment over a size-1 batch (environment/envs/BaseEnvironment.hpp:39-211)
struct EnvBase {
  agarcl_cfg cfg{};
  agarcl_batch* b = nullptr;
  agarcl_layout L{};
  uint64_t seed_value = 0;
  bool seeded = false;
  std::vector<float> dxdy;
  std::vector<int32_t> act;
  std::vector<uint8_t> done;
  void* mirror = nullptr;  // library-owned host copy of the observation
  int64_t oshape[4] = {0, 0, 0, 0};

  EnvBase(int agents, int tps, int arena, bool regen, int pellets, int viruses, int bots, int reward_type, int c_death, int mode,
          int ram_obs) {
    cfg.n_instances = 1;
    cfg.num_agents = agents; cfg.ticks_per_step = tps; cfg.arena_size = arena; cfg.pellet_regen = regen;
    cfg.num_pellets = pellets; cfg.num_viruses = viruses; cfg.num_bots = bots;
    cfg.reward_type = reward_type; cfg.c_death = c_death; cfg.mode_number = mode;
    cfg.num_frames = 1; cfg.grid_size = 128;
    cfg.observe_cells = cfg.observe_others = cfg.observe_viruses = cfg.observe_pellets = 1;
    cfg.rng_mode = AGARCL_RNG_MT19937;  // seed(s) spawns exactly where the reference's mt19937_64 would; unseeded: random_device
    cfg.ram_obs = ram_obs;
    agarcl_layout tmp;
    ck(agarcl_make_layout(&cfg, &tmp));  // the reference constructor throws on a bad mode (Engine::set_mode)
    dxdy.assign((size_t)agents * 2, 0.0f);
    act.assign((size_t)agents, 0);
    done.assign((size_t)agents, 0);
  }
  virtual ~EnvBase() { drop(); }
  void drop() {
    if (b) agarcl_batch_destroy(b);
    b = nullptr;
    mirror = nullptr;
  }
  void ensure() {
    if (b) return;
    ck(agarcl_batch_create(&cfg, &b));
    ck(agarcl_batch_get_layout(b, &L));
    ck(agarcl_batch_reset(b, nullptr, nullptr));  // the reference constructor ends with reset() (BaseEnvironment.hpp:66)
    if (seeded) ck(agarcl_batch_seed(b, &seed_value));
  }
  void configure(const py::dict& d) {  // bindings.cpp:104-114
    auto geti = [&](const char* k, int32_t& v) { if (d.contains(k)) v = d[k].cast<int>(); };
    auto getb = [&](const char* k, int32_t& v) { if (d.contains(k)) v = d[k].cast<bool>() ? 1 : 0; };
    geti("num_frames", cfg.num_frames); geti("grid_size", cfg.grid_size);
    getb("observe_cells", cfg.observe_cells); getb("observe_others", cfg.observe_others);
    getb("observe_viruses", cfg.observe_viruses); getb("observe_pellets", cfg.observe_pellets);
    agarcl_layout tmp;
    ck(agarcl_make_layout(&cfg, &tmp));
    drop();
  }
  void seed(int s) {  // BaseEnvironment::seed (BaseEnvironment.hpp:211)
    seed_value = (uint64_t)(unsigned)s;
    seeded = true;
    if (b) ck(agarcl_batch_seed(b, &seed_value));
  }
  void reset() {
    ensure();
    ck(agarcl_batch_reset(b, nullptr, nullptr));
    ck(agarcl_batch_get_layout(b, &L));  // (strict_reference: the player order of the new episode, quirk Q3)
    std::fill(done.begin(), done.end(), 0);
  }
  void take_actions(const py::list& actions) {  // to_action_vector + take_actions (bindings.cpp:50-64,117-119; BaseEnvironment.hpp:141-176)
    if ((int)actions.size() != cfg.num_agents)
      throw std::runtime_error("Number of actions (" + std::to_string(actions.size()) + ") does not match number of agents (" +
                               std::to_string(cfg.num_agents) + ")");
    size_t a = 0;
    for (auto item : actions) {
      auto t = py::cast<py::tuple>(item);
      dxdy[2 * a] = t[0].cast<float>();
      dxdy[2 * a + 1] = t[1].cast<float>();
      act[a] = t[2].cast<int>();
      a++;
    }
    ensure();
    ck(agarcl_batch_set_actions(b, dxdy.data(), act.data(), 0, nullptr));
  }
  std::vector<double> step() {  // BaseEnvironment::step: rewards in the player map's order of the agents (quirk Q15)
    ensure();
    std::vector<double> rew((size_t)cfg.num_agents, 0.0), out;
    ck(agarcl_batch_step_host(b, dxdy.data(), act.data(), nullptr, rew.data(), done.data()));
    for (int k = 0; k < L.P; k++)
      if (L.order[k] < L.A) out.push_back(rew[(size_t)L.order[k]]);
    return out;
  }
  std::vector<bool> dones() {
    ensure();
    return std::vector<bool>(done.begin(), done.end());
  }
  void save_env_state(const std::string& path) { ensure(); ck(agarcl_batch_save_env_state(b, 0, path.c_str())); }
  void load_env_state(const std::string& path) { ensure(); ck(agarcl_batch_load_env_state(b, 0, path.c_str(), 0)); }
  [[noreturn]] void no_gl() const { throw std::runtime_error("OpenGL rendering is out of scope of agarcl_b200"); }
};

struct GridEnv : EnvBase {
  GridEnv(int agents, int tps, int arena, bool regen, int pellets, int viruses, int bots, int reward_type, int c_death, int mode)
      : EnvBase(agents, tps, arena, regen, pellets, viruses, bots, reward_type, c_death, mode, 0) {}
  py::tuple observation_shape() {
    agarcl_layout tmp;
    ck(agarcl_make_layout(&cfg, &tmp));
    return py::make_tuple(cfg.num_frames * tmp.obs_channels, cfg.grid_size, cfg.grid_size);
  }
  py::list get_state() {  // bindings.cpp:67-91: one fresh int32 (C, G, G) array per agent, owned by numpy
    ensure();
    int32_t dt = 0;
    ck(agarcl_batch_mirror(b, &mirror, oshape, &dt));
    ck(agarcl_batch_sync_mirror(b, nullptr));
    py::list obs;
    const size_t per = (size_t)(oshape[1] * oshape[2] * oshape[3]);
    for (int a = 0; a < cfg.num_agents; a++) {
      py::array_t<int32_t> arr({oshape[1], oshape[2], oshape[3]});
      std::memcpy(arr.mutable_data(), static_cast<const int32_t*>(mirror) + per * (size_t)a, per * sizeof(int32_t));
      obs.append(arr);
    }
    return obs;
  }
};

// ---- GoBigger info structs (environment/envs/GoBiggerEnvironment.hpp:30-243), plain data filled from the records
struct Location { float x = 0, y = 0; };
struct FoodInfo { Location position; float radius = 0; int score = 0; };
struct VirusInfo { Location position; float radius = 0; int score = 0; std::pair<float, float> velocity{0.f, 0.f}; };
struct SporeInfo { Location position; float radius = 0; int score = 0; std::pair<float, float> velocity{0.f, 0.f}; int owner = 0; };
struct CloneInfo { Location position; float radius = 0; int score = 0; std::pair<float, float> velocity{0.f, 0.f}; float direction = 0; int owner = 0; int teamId = 0; };
struct GlobalState {
  int w, h, limit, last, teams;
  GlobalState(int w_, int h_, int limit_, int last_, int teams_) : w(w_), h(h_), limit(limit_), last(last_), teams(teams_) {}
};
struct PlayerState {
  int pid = 0;
  std::vector<FoodInfo> food;
  std::vector<VirusInfo> virus;
  std::vector<SporeInfo> spore;
  std::vector<CloneInfo> clone;
  std::string team;
  double score = 0;
  bool can_eject = true, can_split = true;
};
struct PlayerStates {
  std::unordered_map<int, PlayerState> states;
};

PlayerState from_record(int pid, const float* rec) {
  PlayerState s;
  s.pid = pid;
  const int nf = (int)rec[0], nv = (int)rec[1], ns = (int)rec[2], nc = (int)rec[3];
  s.score = rec[4];
  for (int i = 0; i < nf && i < AGARCL_RAM_KP; i++) {
    const float* e = rec + AGARCL_RAM_OFF_FOOD + 4 * i;
    s.food.push_back(FoodInfo{{e[0], e[1]}, e[2], (int)e[3]});
  }
  for (int i = 0; i < nv && i < AGARCL_RAM_KV; i++) {
    const float* e = rec + AGARCL_RAM_OFF_VIRUS + 4 * i;
    s.virus.push_back(VirusInfo{{e[0], e[1]}, e[2], (int)e[3], {0.f, 0.f}});
  }
  for (int i = 0; i < ns && i < AGARCL_RAM_KS; i++) {
    const float* e = rec + AGARCL_RAM_OFF_SPORE + 4 * i;
    s.spore.push_back(SporeInfo{{e[0], e[1]}, e[2], (int)e[3], {0.f, 0.f}, pid});
  }
  for (int i = 0; i < nc && i < AGARCL_RAM_KC; i++) {
    const float* e = rec + AGARCL_RAM_OFF_CLONE + 8 * i;
    s.clone.push_back(CloneInfo{{e[0], e[1]}, e[2], (int)e[3], {e[4], e[5]}, e[6], (int)e[7], 0});
  }
  return s;
}

struct GoBiggerEnv : EnvBase {
  GlobalState global;
  int frames = 0;
  GoBiggerEnv(int map_w, int map_h, int frame_limit, int agents, int tps, int arena, bool regen, int pellets, int viruses, int bots,
              int reward_type, int c_death, int mode, bool /*load_env_snapshot*/, bool /*agent_view*/)
      : EnvBase(agents, tps, arena, regen, pellets, viruses, bots, reward_type, c_death, mode, 1), global(map_w, map_h, frame_limit, 0, agents) {}
  void reset_() { reset(); frames = 0; }
  std::vector<double> step_() {
    auto r = step();
    frames += cfg.num_agents;  // one add_frame per agent per step (GoBiggerEnvironment.hpp:515-521)
    return r;
  }
  py::tuple observation_shape() const { return py::make_tuple(frames, global.h, global.w); }
  py::list get_state() {  // bindings.cpp:28-47
    ensure();
    std::vector<float> ram((size_t)L.P * AGARCL_RAM_RECORD);
    ck(agarcl_batch_ram_host(b, ram.data()));
    PlayerStates ps;
    for (int p = 0; p < L.P; p++) {
      const float* rec = ram.data() + (size_t)p * AGARCL_RAM_RECORD;
      if (rec[0] + rec[1] + rec[2] + rec[3] > 0) ps.states.emplace(p, from_record(p, rec));
    }
    py::dict d;
    d["global_state"] = global;
    d["player_states"] = ps;
    py::list out;
    out.append(d);
    return out;
  }
};

template <class T>
void bind_position(py::class_<T>& c) {
  c.def_readwrite("position", &T::position)
      .def("get_position_x", [](const T& f) { return f.position.x; })
      .def("get_position_y", [](const T& f) { return f.position.y; })
      .def_readwrite("radius", &T::radius)
      .def_readwrite("score", &T::score);
}

}  // namespace

PYBIND11_MODULE(agarcl, m) {
  m.doc() = "agarcl: the reference's Python module (environment/bindings.cpp) over the B200 CUDA library libagarcl_b200.so";

  py::class_<GridEnv>(m, "GridEnvironment")  // bindings.cpp:99-135
      .def(py::init<int, int, int, bool, int, int, int, int, int, int>())
      .def("seed", &GridEnv::seed)
      .def("configure_observation", &GridEnv::configure)
      .def("observation_shape", &GridEnv::observation_shape)
      .def("dones", &GridEnv::dones)
      .def("take_actions", &GridEnv::take_actions)
      .def("reset", &GridEnv::reset)
      .def("render", [](GridEnv& e) { e.no_gl(); })
      .def("step", &GridEnv::step)
      .def("get_state", &GridEnv::get_state)
      .def("get_frame", [](GridEnv& e) { e.no_gl(); })
      .def("close", [](GridEnv& e) { e.drop(); })
      .def("save_env_state", &GridEnv::save_env_state);
  m.attr("has_screen_env") = py::bool_(false);  // bindings.cpp:172-176

  py::class_<Location>(m, "Location").def_readwrite("x", &Location::x).def_readwrite("y", &Location::y);
  { py::class_<FoodInfo> c(m, "FoodInfo"); bind_position(c); }
  { py::class_<VirusInfo> c(m, "VirusInfo"); bind_position(c); c.def_readwrite("velocity", &VirusInfo::velocity); }
  { py::class_<SporeInfo> c(m, "SporeInfo"); bind_position(c); c.def_readwrite("velocity", &SporeInfo::velocity).def_readwrite("owner", &SporeInfo::owner); }
  { py::class_<CloneInfo> c(m, "CloneInfo"); bind_position(c);
    c.def_readwrite("velocity", &CloneInfo::velocity).def_readwrite("direction", &CloneInfo::direction)
        .def_readwrite("owner", &CloneInfo::owner).def_readwrite("teamId", &CloneInfo::teamId); }
  py::class_<GlobalState>(m, "GlobalState")  // bindings.cpp:228-247
      .def(py::init<int, int, int, int, int>(), py::arg("width"), py::arg("height"), py::arg("frame_limit"), py::arg("last_frame"), py::arg("team_num"))
      .def("update_last_frame_count", [](GlobalState& g, int n) { g.last = n; })
      .def("get_map_width", [](const GlobalState& g) { return g.w; })
      .def("get_map_height", [](const GlobalState& g) { return g.h; })
      .def("get_frame_limit", [](const GlobalState& g) { return g.limit; })
      .def("get_team_num", [](const GlobalState& g) { return g.teams; })
      .def("__str__", [](const GlobalState& g) {
        std::ostringstream o;
        o << "GlobalState(map_width=" << g.w << ", map_height=" << g.h << ", frame_limit=" << g.limit << ", team_num=" << g.teams << ")";
        return o.str();
      });
  py::class_<PlayerState>(m, "PlayerState")  // bindings.cpp:250-277
      .def("get_player_id", [](const PlayerState& s) { return s.pid; })
      .def("get_food_infos", [](const PlayerState& s) { return s.food; })
      .def("get_virus_infos", [](const PlayerState& s) { return s.virus; })
      .def("get_spore_infos", [](const PlayerState& s) { return s.spore; })
      .def("get_clone_infos", [](const PlayerState& s) { return s.clone; })
      .def("get_team_name", [](const PlayerState& s) { return s.team; })
      .def("get_score", [](const PlayerState& s) { return s.score; })
      .def("canEject", [](const PlayerState& s) { return s.can_eject; })
      .def("canSplit", [](const PlayerState& s) { return s.can_split; })
      .def("update_score", [](PlayerState& s, double v) { s.score = v; });
  py::class_<PlayerStates>(m, "PlayerStates")  // bindings.cpp:280-301
      .def("get_player_state", [](PlayerStates& p, int pid) -> PlayerState& {
        auto it = p.states.find(pid);
        if (it == p.states.end()) { PlayerState s; s.pid = pid; it = p.states.emplace(pid, s).first; }
        return it->second;
      }, py::return_value_policy::reference_internal)
      .def("get_all_player_states", [](const PlayerStates& p) { return p.states; })
      .def("__str__", [](const PlayerStates& p) {
        std::ostringstream o;
        o << "PlayerStates:\n";
        for (const auto& kv : p.states)
          o << "  Player " << kv.second.pid << ": score=" << kv.second.score << ", food_seen=" << kv.second.food.size()
            << ", virus_seen=" << kv.second.virus.size() << ", spores_seen=" << kv.second.spore.size()
            << ", no_clone=" << kv.second.clone.size() << ", team_name=\"" << kv.second.team << "\"\n";
        return o.str();
      });

  py::class_<GoBiggerEnv>(m, "GoBiggerEnvironment")  // bindings.cpp:323-374
      .def(py::init<int, int, int, int, int, int, bool, int, int, int, int, int, int, bool, bool>(), py::arg("map_width"), py::arg("map_height"),
           py::arg("frame_limit"), py::arg("num_agents"), py::arg("ticks_per_step"), py::arg("arena_size"), py::arg("pellet_regen"),
           py::arg("num_pellets"), py::arg("num_viruses"), py::arg("num_bots"), py::arg("reward_type"), py::arg("c_death") = 0,
           py::arg("mode_number") = 0, py::arg("load_env_snapshot") = false, py::arg("agent_view") = false)
      .def("configure_observation", &GoBiggerEnv::configure)
      .def("get_state", &GoBiggerEnv::get_state)
      .def("get_frame", [](GoBiggerEnv& e) { e.no_gl(); })
      .def("take_actions", &GoBiggerEnv::take_actions)
      .def("dones", &GoBiggerEnv::dones)
      .def("observation_shape", &GoBiggerEnv::observation_shape)
      .def("seed", &GoBiggerEnv::seed, "Seed the environment")
      .def("reset", &GoBiggerEnv::reset_, "Reset the environment")
      .def("step", &GoBiggerEnv::step_, "Step through the environment")
      .def("render", [](GoBiggerEnv& e) { e.no_gl(); }, "Render the current state")
      .def("close", [](GoBiggerEnv& e) { e.drop(); }, "Close the environment")
      .def("load_env_state", &GoBiggerEnv::load_env_state)
      .def("save_env_state", &GoBiggerEnv::save_env_state);
}
