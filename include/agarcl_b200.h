/*
 * agarcl_b200.h — C-ABI of the B200-native batched AgarCL simulator.
 *
 * This is the drop-in boundary for ONE hot path of machado-research/AgarCL:
 *   N lockstep instances of  BaseEnvironment::step  (environment/envs/BaseEnvironment.hpp:89-122)
 *     = ticks_per_step x Engine::tick (agario/engine/Engine.hpp:208-240), built-in bots
 *       (agario/bots/ headers), regen/respawn, rewards, dones,
 *   + GridObservation::add_frame (environment/envs/GridEnvironment.hpp:91-123).
 *
 * Every entry point below replaces what the reference's pybind11 module `agarcl`
 * (environment/bindings.cpp:94-135) binds for that path; the cited line is the
 * reference interface it stands in for.  Plain pointers and sizes only: no C++
 * or torch types cross this boundary.  All functions return 0 on success and a
 * negative agarcl_status on failure; agarcl_last_error() gives the text.
 *
 * The same header also fixes the per-instance STATE BLOB layout that is shared by
 *   - the CUDA library (device state = N blobs, `stride` bytes apart),
 *   - oracle/oracle.c (CPU restatement, test infrastructure only),
 *   - oracle/ref_harness.cpp (the compiled reference, dumps into the same blob).
 */
#ifndef AGARCL_B200_H
#define AGARCL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- constants
 * agario/core/settings.hpp, agario/core/Entities.hpp.     */
#define AGARCL_CELL_MIN_SIZE 25u
#define AGARCL_CELL_SPLIT_MINIMUM 50u
#define AGARCL_PLAYER_CELL_LIMIT 14
#define AGARCL_FOOD_MASS 10u
#define AGARCL_PELLET_MASS 1u
#define AGARCL_VIRUS_INITIAL_MASS 100u
#define AGARCL_MAX_MASS_IN_THE_GAME 22500u
#define AGARCL_NEW_MASS_IF_NO_SPLIT 22000u
#define AGARCL_RECOMBINE_TICKS 300u /* RECOMBINE_TIMER_SEC(10) * 30 ticks/s, sim-time */

#define AGARCL_MAX_CELLS 32   /* cell slots per player (reference limit 14, Engine.hpp:592-601 can overshoot) */
#define AGARCL_VET_CAP 16     /* remembered virus_eaten_ticks per player (Player.hpp:31) */
#define AGARCL_MAX_PLAYERS 64 /* agents + bots per instance */
#define AGARCL_LUT_SIZE 65536 /* integer-mass lookup tables (radius, max_speed, split_speed) */

/* hdr.flags bits: set by the simulator when a fixed capacity was hit or a
 * non-replayable reference path was taken.  Parity is only claimed while 0.  */
#define AGARCL_FLAG_FOOD_OVERFLOW 0x001u
#define AGARCL_FLAG_VIRUS_OVERFLOW 0x002u
#define AGARCL_FLAG_CELL_OVERFLOW 0x004u
#define AGARCL_FLAG_VET_OVERFLOW 0x008u
#define AGARCL_FLAG_EATER_OVERFLOW 0x010u
#define AGARCL_FLAG_REPLAY_EXHAUSTED 0x020u
#define AGARCL_FLAG_PCD_TIE 0x040u     /* (round 1: >16 cells in one PCD strip with equal y.  No longer raised: libstdc++'s std::sort is restated exactly, ties included) */
#define AGARCL_FLAG_RAND_SITE 0x080u   /* a libc rand() site was reached (Bot.hpp:93-96,121-126) */
#define AGARCL_FLAG_MASS_LUT 0x100u    /* a cell mass exceeded AGARCL_LUT_SIZE */
#define AGARCL_FLAG_REMOVE_OVERFLOW 0x200u

typedef enum agarcl_status {
  AGARCL_OK = 0,
  AGARCL_ERR_INVALID = -1,   /* bad argument / config (EnvironmentException in the reference) */
  AGARCL_ERR_CUDA = -2,      /* CUDA runtime failure */
  AGARCL_ERR_NOMEM = -3,
  AGARCL_ERR_STATE = -4      /* call order (e.g. step before reset) */
} agarcl_status;

typedef enum agarcl_rng_mode {
  AGARCL_RNG_PHILOX = 0,  /* counter-based per-instance Philox4x32-10 on device */
  AGARCL_RNG_REPLAY = 1,  /* draws come from a per-instance stream uploaded with agarcl_batch_set_replay */
  AGARCL_RNG_MT19937 = 2  /* the host keeps one std::mt19937_64 per instance -- seeded by agarcl_batch_seed exactly as
                             Engine::seed does (Engine.hpp:242-245), from std::random_device when never seeded
                             (GameState.hpp:59) -- and feeds its uniform_real_distribution<float> draws
                             (random_location, Engine.hpp:143-148) to the device through a ring that it refills
                             ahead of every instance's cursor: the stream never runs out */
} agarcl_rng_mode;

typedef enum agarcl_obs_dtype { AGARCL_OBS_I32 = 0, AGARCL_OBS_I16 = 1 } agarcl_obs_dtype;

/* ------------------------------------------------------------------- config
 * Fields 2..11 are the positional arguments of agarcl.GridEnvironment
 * (environment/bindings.cpp:102, BaseEnvironment.hpp:39-51); fields 12..17 are
 * the keys of configure_observation (bindings.cpp:104-114).                   */
typedef struct agarcl_cfg {
  int32_t n_instances;
  int32_t num_agents, ticks_per_step, arena_size, pellet_regen, num_pellets, num_viruses, num_bots;
  int32_t reward_type, c_death, mode_number;
  int32_t num_frames, grid_size, observe_cells, observe_others, observe_viruses, observe_pellets;
  int32_t obs_dtype;        /* agarcl_obs_dtype; int32 is the reference's */
  int32_t strict_reference; /* 1: keep quirk Q11 (frame index) exactly, and quirk Q3: the player order of the episode that follows the
                               k-th unmasked agarcl_batch_reset is the one the reference has after its k-th reset (its player map and pid
                               counter outlive a reset, Engine.hpp:72,98-101); 0: always render the last frame(s), every episode in the
                               player order of a fresh engine */
  int32_t rng_mode;         /* agarcl_rng_mode */
  int32_t cap_viruses, cap_foods, cap_replay; /* 0 = defaults */
  int32_t device;           /* CUDA device ordinal */
  int32_t instance_base;    /* global index of local instance 0 (multi-GPU sharding; keys the RNG) */
  int32_t ram_obs;          /* 1: every step also produces the structured observation (agarcl_batch_ram); 2: ONLY that one
                               (agario-ram-v0: the grid frame is not rendered, agarcl_batch_obs stays zero) */
  int32_t reserved[2];
} agarcl_cfg;

/* ------------------------------------------------------------- state records */
typedef struct agarcl_cell { /* 48 B; agario/core/Entities.hpp:118-212 */
  float x, y, vx, vy;
  float svx, svy;        /* splitting_velocity */
  uint32_t mass;
  uint32_t id;           /* creation order inside the instance (Ball.hpp:15,18; only relative order is used) */
  uint32_t recomb_tick;  /* engine tick from which can_recombine() holds (Entities.hpp:183-193, sim-time) */
  uint32_t pad[3];
} agarcl_cell;

typedef struct agarcl_virus { /* 32 B; Entities.hpp:82-114 */
  float x, y;
  uint32_t mass;
  int32_t hits;          /* _num_food_hits */
  float vx, vy;
  uint32_t pad[2];
} agarcl_virus;

typedef struct agarcl_food { float x, y, vx, vy; } agarcl_food; /* 16 B; Entities.hpp:41-55 */
typedef struct agarcl_pellet { float x, y; } agarcl_pellet;     /* 8 B;  Entities.hpp:22-38 */

typedef struct agarcl_player { /* 128 B; agario/core/Player.hpp:25-41,211-215 */
  int32_t n_cells;
  float target_x, target_y;
  int32_t action;          /* 0 none, 1 feed, 2 split (core/types.hpp:59-61) */
  int32_t split_cd, feed_cd;
  float anti_team_decay;
  int32_t elapsed_ticks, last_decay_tick;
  int32_t bot_type;        /* -1 agent; 0 Hungry, 1 HungryShy, 2 Aggressive, 3 AggressiveShy */
  uint32_t min_mass_cell;
  int32_t food_eaten;
  uint32_t highest_mass;
  int32_t cells_eaten, viruses_eaten;
  int32_t vet_count;       /* virus_eaten_ticks.size() */
  int32_t vet_ticks[AGARCL_VET_CAP];
} agarcl_player;

typedef struct agarcl_inst_hdr { /* 64 B */
  uint32_t tick;           /* GameState::ticks */
  uint32_t next_cell_id;
  int32_t n_pellets, n_viruses, n_foods;
  uint32_t rng_cursor;     /* uniform draws consumed so far */
  uint32_t flags;          /* AGARCL_FLAG_* */
  uint32_t seed_lo, seed_hi;
  uint32_t done_sticky;    /* mode 3 sticky done (BaseEnvironment.hpp:132-135) */
  uint32_t respawned_lo, respawned_hi; /* players respawned at the end of the last step: the step's
                             observation is taken BEFORE repsawn_all_players (BaseEnvironment.hpp:96-101) */
  uint32_t pad[4];
} agarcl_inst_hdr;

/* Byte offsets of one instance's arrays inside its blob. */
typedef struct agarcl_layout {
  int32_t P, A;            /* players (agents + bots), agents */
  int32_t cap_cells, cap_pellets, cap_viruses, cap_foods, cap_replay;
  uint32_t off_hdr, off_players, off_cells, off_viruses, off_foods, off_pellets;
  uint32_t stride;         /* bytes per instance, multiple of 128 */
  /* engine mode flags, Engine::set_mode (Engine.hpp:367-416) */
  int32_t mass_decay, squared_pellets, regen, agent_mass;
  int32_t obs_channels;    /* per frame: 1 + cells + 2*others + 2*viruses + 2*pellets (GridEnvironment.hpp:188-195) */
  int32_t order[AGARCL_MAX_PLAYERS];    /* player index processed k-th (libstdc++ unordered_map order) */
  int32_t bot_type[AGARCL_MAX_PLAYERS];
} agarcl_layout;

/* Fills `L` from `c` (host side; uses the same libstdc++ unordered_map the reference
 * iterates, GameState.hpp:44).  Returns AGARCL_ERR_INVALID on a bad config.  */
int agarcl_make_layout(const agarcl_cfg* c, agarcl_layout* L);

/* ---------------------------------------------------------------- batch API */
typedef struct agarcl_batch agarcl_batch;

/* GridEnvironment ctor + configure_observation (bindings.cpp:102-114). Allocates all device memory. */
int agarcl_batch_create(const agarcl_cfg* cfg, agarcl_batch** out);
int agarcl_batch_destroy(agarcl_batch* b);
int agarcl_batch_get_layout(const agarcl_batch* b, agarcl_layout* out);

/* BaseEnvironment::seed (BaseEnvironment.hpp:211): seeds[i] for instance i (host pointer, N entries). */
int agarcl_batch_seed(agarcl_batch* b, const uint64_t* seeds);
/* BaseEnvironment::reset (BaseEnvironment.hpp:179-204) for instances with mask[i]!=0 (NULL = all).  Like the reference's,
 * a reset does not reseed: the instance's draw stream goes on where the last episode left it (successive episodes
 * differ); only agarcl_batch_seed / agarcl_batch_set_replay / a snapshot load restart it at its first draw. */
int agarcl_batch_reset(agarcl_batch* b, const uint8_t* mask, void* stream);
/* (strict_reference: an unmasked reset changes layout.order -- fetch it again with agarcl_batch_get_layout) */
/* BaseEnvironment::take_actions (BaseEnvironment.hpp:141-176): dxdy[N*A*2], act[N*A].
 * on_device!=0: device pointers, read by the next step without a copy. */
int agarcl_batch_set_actions(agarcl_batch* b, const float* dxdy, const int32_t* act, int on_device, void* stream);
/* BaseEnvironment::step + GridObservation::add_frame for every agent, enqueued on `stream` (cudaStream_t). */
int agarcl_batch_step(agarcl_batch* b, void* stream);
/* get_state (bindings.cpp:67-91) without the copy: device pointer to obs [N*A, C*num_frames, G, G]. */
int agarcl_batch_obs(agarcl_batch* b, void** dev_ptr, int64_t shape[4], int32_t* dtype);
int agarcl_batch_rewards(agarcl_batch* b, double** dev_ptr); /* step() return value, f64[N*A] */
int agarcl_batch_dones(agarcl_batch* b, uint8_t** dev_ptr);  /* dones(), u8[N*A] */
/* The reference-facing call with HOST buffers: copies actions in, steps, copies obs/rewards/dones
 * out (any of the out pointers may be NULL) and synchronises.  This is what `e2e` in bench.py times. */
int agarcl_batch_step_host(agarcl_batch* b, const float* dxdy, const int32_t* act,
                           void* obs_out, double* rewards_out, uint8_t* dones_out);

/* ---------------------------------------------------- host-resident observation mirror
 * get_state (bindings.cpp:67-91) hands every agent's frame to HOST memory; for a batch that is 512 KB per agent
 * per step, and the dense copy above is bound by PCIe.  The mirror is a library-owned, pinned host tensor
 * [N*A, C*num_frames, G, G] of the observation dtype that the library keeps IDENTICAL to the device observation
 * while moving only what a grid observation really contains: per frame the row/column bit masks of the
 * out-of-bounds channel (GridEnvironment.hpp:235-248) and the (offset, value) list of the few hundred non-zero
 * elements of the other channels (:212-232).  Host threads undo the previous step's list, patch the mask rows /
 * columns that changed and store the new list.  Images that do not fit the scheme are copied densely, so the
 * result never depends on it.  The caller reads the mirror and must not write to it.
 *   agarcl_batch_mirror       creates the mirror on first use, syncs it, returns the host pointer (stable).
 *   agarcl_batch_sync_mirror  brings it up to date with the current device observation (after a reset, a
 *                             render, or agarcl_batch_step); synchronises `stream`.
 *   agarcl_batch_step_mirror  = take_actions (host) + step + sync_mirror + rewards/dones to host: the
 *                             reference-facing call `e2e` in bench.py times.  When the step is the single fused
 *                             kernel (int32, one frame) that kernel lists what it scatters itself and flags
 *                             finished chunks of instances in host-mapped memory, so the host fetches and
 *                             expands chunk k while the device still steps chunk k+1.
 *   agarcl_batch_mirror_stats out[0] entries moved by the last sync, [1] images copied densely, [2] bytes copied
 *                             device->host, [3] host threads.                                                  */
int agarcl_batch_mirror(agarcl_batch* b, void** host_ptr, int64_t shape[4], int32_t* dtype);
int agarcl_batch_sync_mirror(agarcl_batch* b, void* stream);
int agarcl_batch_step_mirror(agarcl_batch* b, const float* dxdy, const int32_t* act, double* rewards_out, uint8_t* dones_out);
int agarcl_batch_mirror_stats(const agarcl_batch* b, uint64_t out[4]);
/* last agarcl_batch_step_mirror, microseconds: out[0] the calling thread waiting for the device (chunk flags, list
 * copies), out[1] first wait to mirror complete, out[2] action staging + kernel launch, out[3] the whole call. */
int agarcl_batch_mirror_timing(const agarcl_batch* b, uint64_t out[4]);

/* ---------------------------------------------------- the observation as LISTS (zero-copy, no dense host tensor)
 * A grid observation is almost entirely structure: per frame the out-of-bounds channel is a row mask and a column mask
 * (GridEnvironment.hpp:235-248) and the other channels hold a few dozen non-zero elements (:212-232).  The step kernel
 * writes exactly that -- per image a record and a list of integer operations -- straight into pinned host memory while it
 * runs; the dense mirror above is those lists replayed by host threads, which is what stops scaling when 8 GPUs feed one
 * host.  agarcl_batch_step_lists hands the lists themselves to the caller: take_actions (host) + step +
 * rewards / dones to host, observation left in library-owned pinned memory as described by agarcl_obs_lists, valid until
 * the next step call on the batch.  A learner expands them where it trains (one scatter per image) or consumes them as
 * they are; agarcl_batch_lists_expand is the reference decoder for one image.  Requires the configuration whose step is
 * the single fused kernel (one frame, strict_reference = 0); AGARCL_ERR_STATE otherwise.
 *   chunk c                = chunks + c * chunk_words                     (32-bit words)
 *   slot s (of n_images)   : chunk s / images_per_chunk, index li = s % images_per_chunk inside it; slot_of[image] = s
 *   record of the slot     = chunk + off_rec + li * rec_words :
 *        [0] entries (0xFFFFFFFF: the image overflowed entries_per_image -> read it from agarcl_batch_obs)
 *        [1] index of its first entry in the chunk's entry array   [2] the image (instance * agents + agent)
 *        [3 ..] per frame: mask_words words of ROW bits (bit i: row i of channel 0 is -1), then mask_words of COLUMN bits
 *        [rec_words-3, -2] the agent's reward (f64)   [rec_words-1] its done flag
 *   entries of the chunk   = (uint32 pairs) chunk + off_entries : (op << 29 | element offset inside the image, operand), applied in
 *        order onto a zero frame: op 0 x = v, 1 x += v, 2 x = (x != 0 && x < v) ? x : v, 3 x = max(x, v); int16 saturates at 32767 */
typedef struct agarcl_obs_lists {
  int32_t n_images, n_chunks, images_per_chunk, frames, channels, grid, obs_dtype;
  int32_t mask_words, rec_words, entries_per_image;
  uint32_t off_rec, off_entries, chunk_words;
  const uint32_t* chunks;   /* pinned host memory */
  const uint32_t* slot_of;  /* [n_images] */
} agarcl_obs_lists;
int agarcl_batch_step_lists(agarcl_batch* b, const float* dxdy, const int32_t* act, double* rewards_out, uint8_t* dones_out,
                            agarcl_obs_lists* out);
/* Decodes image `image` of the lists of the last agarcl_batch_step_lists into dense_out ([frames*C, G, G] of the observation
 * dtype, host memory); an image that overflowed its slot is copied from the device tensor instead. */
int agarcl_batch_lists_expand(agarcl_batch* b, int32_t image, void* dense_out);

/* ---------------------------------------------------- structured ("ram") observation
 * GoBiggerObservation::add_frame (environment/envs/GoBiggerEnvironment.hpp:515-548, _store_entities
 * 446-513): for EVERY player of the instance, the in-view viruses, pellets ("food"), ejected foods
 * ("spores") and the player's own cells ("clones"), player-relative, in entity index order — what
 * agarcl.GoBiggerEnvironment.get_state() (environment/bindings.cpp:28-47,323-374) hands out as
 * PlayerState objects.  Here one player is one fixed-size float32 record, zero padded:
 *   hdr   [8]      n_food, n_virus, n_spore, n_clone (true in-view counts), score (= player mass),
 *                  player x, player y, overflow mask (bit0 food, bit1 virus, bit2 spore, bit3 clone)
 *   food  [KP][4]  dx, dy, radius, score            (FoodInfo)
 *   virus [KV][4]  dx, dy, radius, score            (VirusInfo; its velocity is the constant (0,0))
 *   spore [KS][4]  dx, dy, radius, score            (SporeInfo; velocity (0,0), owner = this player)
 *   clone [KC][8]  dx, dy, radius, score, vx, vy, direction, owner   (CloneInfo; teamId is always 0)
 * A player with nothing in view (a dead one) keeps its previous record, like the reference, whose
 * PlayerState is only committed when an entity lands inside the grid (:495-509).  reset() clears.   */
#define AGARCL_RAM_HDR 8
#define AGARCL_RAM_KP 192
#define AGARCL_RAM_KV 16
#define AGARCL_RAM_KS 32
#define AGARCL_RAM_KC 32
#define AGARCL_RAM_OFF_FOOD AGARCL_RAM_HDR
#define AGARCL_RAM_OFF_VIRUS (AGARCL_RAM_OFF_FOOD + 4 * AGARCL_RAM_KP)
#define AGARCL_RAM_OFF_SPORE (AGARCL_RAM_OFF_VIRUS + 4 * AGARCL_RAM_KV)
#define AGARCL_RAM_OFF_CLONE (AGARCL_RAM_OFF_SPORE + 4 * AGARCL_RAM_KS)
#define AGARCL_RAM_RECORD (AGARCL_RAM_OFF_CLONE + 8 * AGARCL_RAM_KC) /* 1224 floats per player */
/* Device pointer to the records [N, P, AGARCL_RAM_RECORD] float32 (cfg.ram_obs must be 1). */
int agarcl_batch_ram(agarcl_batch* b, float** dev_ptr, int64_t shape[3]);
/* get_state() of agarcl.GoBiggerEnvironment (bindings.cpp:28-47) with a HOST buffer: all records [N, P, AGARCL_RAM_RECORD]
 * copied out (synchronises the device). */
int agarcl_batch_ram_host(agarcl_batch* b, float* out);
/* Run only the structured-observation kernel on the current state (tests). */
int agarcl_batch_render_ram(agarcl_batch* b, void* stream);

/* Parity / snapshot transport: one instance's blob (layout.stride bytes) to / from host memory. */
int agarcl_batch_download_state(agarcl_batch* b, int32_t instance, void* blob);
int agarcl_batch_upload_state(agarcl_batch* b, int32_t instance, const void* blob);
/* BaseEnvironment::save_env_state / load_env_state (environment/envs/BaseEnvironment.hpp:213-343,
 * agario/engine/Engine.hpp:247-348; bound at bindings.cpp:135,170,374): one instance to / from the
 * reference's JSON snapshot.  The reference format is lossy (no splitting velocity, recombine timers, virus
 * food hits, tick): those travel as extra keys the reference ignores and are restored with lossless != 0.
 * After a load the instance's draw stream restarts from the snapshot's seed, like Engine::seed there.   */
int agarcl_batch_save_env_state(agarcl_batch* b, int32_t instance, const char* path);
int agarcl_batch_load_env_state(agarcl_batch* b, int32_t instance, const char* path, int lossless);
/* The same on a host blob of `L->stride` bytes (no device needed): what the two calls above run. */
int agarcl_snapshot_write(const agarcl_cfg* c, const agarcl_layout* L, const void* blob, const char* path);
int agarcl_snapshot_read(const agarcl_cfg* c, const agarcl_layout* L, void* blob, const char* path, int lossless);
/* Recorded uniform draws in [0,1) for instance i (replay of the reference's mt19937_64). */
int agarcl_batch_set_replay(agarcl_batch* b, int32_t instance, const float* draws, int32_t n);
/* Run only the observation kernel on the current state (tests; add_frame on a cleared buffer). */
int agarcl_batch_render(agarcl_batch* b, void* stream);
/* Per-kernel device timing (bench.py roofline): when enabled, every agarcl_batch_step brackets the
 * engine-tick kernel and the observation kernel with CUDA events on the launching stream.
 * agarcl_batch_get_timing synchronises, returns the accumulated milliseconds and step count since the
 * last call, and resets the accumulators. */
int agarcl_batch_set_timing(agarcl_batch* b, int enable);
int agarcl_batch_get_timing(agarcl_batch* b, double* sim_ms, double* obs_ms, int32_t* steps);
/* Number of kernel launches issued by the last agarcl_batch_step. */
int agarcl_batch_launches_per_step(const agarcl_batch* b);
/* hdr.flags of the whole batch, reduced on the device: *or_all = OR over all instances, counts[bit] = instances with
 * AGARCL_FLAG bit `bit` set (either pointer may be NULL).  Synchronises `stream`.  The reference has no counterpart: its
 * containers grow without bound where this library has fixed capacities (BaseEnvironment.hpp / Engine.hpp vectors). */
int agarcl_batch_flags(agarcl_batch* b, void* stream, uint32_t* or_all, uint32_t counts[32]);
/* Diagnostics of the schedule: cycles every instance worked in the last step (what the cost-sorted schedule orders the next step by;
 * waiting at the alignment barriers excluded).  out: HOST array of N entries.  Synchronises `stream`.  No reference counterpart. */
int agarcl_batch_costs(agarcl_batch* b, void* stream, uint32_t* out);
/* Self-test of the device's restatement of libstdc++'s std::sort (k_step's strip_std_sort; the reference sorts the strips of
 * PrecisionCollisionDetection::solve with it, agario/utils/collision_detection.hpp:29-31): sorts the indices 0..n-1 by the HOST keys
 * ys[n] on the current device and writes the resulting order to order_out[n].  n <= 65535.  tests/test_gpu_std_sort.py compares it
 * with the oracle's restatement, which is pinned against the real std::sort. */
int agarcl_selftest_std_sort(const float* ys, int32_t n, uint16_t* order_out);
/* Host helper: first n canonical floats of std::mt19937_64(seed) as uniform_real_distribution<float>
 * draws them (random.hpp:6-20). */
int agarcl_mt19937_draws(uint64_t seed, float* out, int32_t n);

const char* agarcl_last_error(void);
const char* agarcl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AGARCL_B200_H */
