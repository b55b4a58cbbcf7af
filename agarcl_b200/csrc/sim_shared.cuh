// sim_shared.cuh — shared-memory carve-up of one warp (= one game instance) in k_step.
//
// The byte offsets of the arrays are computed once on the host (SmemOff, part of SimParams, i.e. the
// constant bank); on the device an array pointer is `warp base + constant`, formed where it is used,
// so the carve-up costs no live registers on the hot path.
#pragma once
#include <cstdint>

#include "../../include/agarcl_b200.h"
#include "sim_params.h"

namespace ag {

__host__ __device__ inline uint32_t ag_align16(uint32_t x) { return (x + 15u) & ~15u; }

// Fills the offsets; returns the bytes one warp needs.  Three groups of arrays share bytes because they are
// never live at the same time (16 instances per SM instead of 13 at configs[1]):
//   * the player-loop scratch (cand, prem, vrem, lprem) and the snapshot of the players_collision pre-test (snap):
//     the removal lists are consumed by apply_removals before players_collision starts;
//   * the exact collision sweep (rows, strip, pairs, reskeys, resorder) lives in the bytes of the pellet hash's index
//     array when that is large enough: the sweep is rare (something must be edible) and simply invalidates the hash,
//     which is rebuilt at the start of the next tick;
//   * the 32-bit counters of the hash build (htmp) lie over cellref + scratch + hitq, all dead at the start of a tick.
inline uint32_t make_smem_offsets(const agarcl_layout& L, int HG, SmemOff& o) {
  uint32_t p = 0;
  // pellet spatial hash: persists across the ticks of a launch, patched on removals
  o.hcnt = p;    p += ag_align16((uint32_t)(HG * HG) * 2u);        // u16 [HG*HG]   cell ends (start of cell k = end of cell k-1)
  o.hsorted = p; p += ag_align16((uint32_t)L.cap_pellets * 2u);    // u16 [cap_pellets] pellet indices grouped by hash cell
  const uint32_t sweep_bytes = 2u * ag_align16(kCellRefCap * 2u) + kPairCap * 16u + 2u * ag_align16(kPairCap * 2u);
  const bool sweep_in_hash = ag_align16((uint32_t)L.cap_pellets * 2u) >= sweep_bytes;
  uint32_t sw = sweep_in_hash ? o.hsorted : p;
  if (!sweep_in_hash) p += sweep_bytes;
  o.rows = sw;     sw += ag_align16(kCellRefCap * 2u);             // i16 [kCellRefCap] strip id
  o.strip = sw;    sw += ag_align16(kCellRefCap * 2u);             // u16 [kCellRefCap] one strip sorted by y
  o.pairs = sw;    sw += kPairCap * 16u;                           // uint4 [kPairCap] PairRec
  o.reskeys = sw;  sw += ag_align16(kPairCap * 2u);                // u16 [kPairCap]
  o.resorder = sw;                                                 // u16 [kPairCap]
  o.sweep_in_hash = sweep_in_hash ? 1u : 0u;
  // the instance's pellet array itself (index order), brought in by ONE TMA bulk load when the warp takes the
  // instance and written back by one bulk store if a pellet was eaten or spawned: every pellet access of the
  // ticks and of the observation scatter is a shared-memory access, not a trip to L2 / HBM behind the obs stores
  o.spel = p;    p += ag_align16((uint32_t)L.cap_pellets * 8u);    // float2 [cap_pellets]
  o.mbar = p;    p += 16u;                                         // u64 mbarrier of the bulk load
  o.cold = p;    p += kColdCtxBytes;                               // ColdCtx (sim_kernel.cu): the instance's warp-uniform, rarely used state
  // players_collision snapshot enumeration (also the y-mask row of the fused observation finish)
  const uint32_t tmp0 = p;
  o.cellref = p; p += ag_align16(kCellRefCap * 2u);                // u16 [kCellRefCap] (player << 8 | cell) in snapshot order
  // player-loop scratch | pre-test snapshot
  const uint32_t u0 = p;
  o.cand = p;    p += ag_align16(kCandCap * 8u);                   // uint2 [kCandCap] (order key, d^2 bits)
  o.prem = p;    p += ag_align16(kPremCap * 2u);                   // u16 [kPremCap] pellets_to_remove
  o.vrem = p;    p += ag_align16(kVremCap * 2u);                   // u16 [kVremCap] viruses_to_remove
  o.lprem = p;   p += ag_align16(32u * kLaneCand * 2u);            // u16 [32][kLaneCand] pellets eaten by each lane's player
  o.snap = u0;                                                     // f32 x[kSnapCap], f32 y[kSnapCap], u32 (mass | player << 24)[kSnapCap]
  if (p - u0 < kSnapCap * 12u) p = u0 + kSnapCap * 12u;
  o.hitq = p;    p += ag_align16(kPairCap * 2u);                   // u16 [kPairCap] queries flagged by the pre-test
  o.htmp = tmp0;                                                   // u32 [HG*HG] counters of the hash build
  if (p - tmp0 < (uint32_t)(HG * HG) * 4u) p = tmp0 + ag_align16((uint32_t)(HG * HG) * 4u);
  // live across ticks
  o.vcache = p;  p += ag_align16((uint32_t)L.cap_viruses * 16u);   // float4 [cap_viruses] x, y, radius, mass bits
  o.psum = p;    p += ag_align16((uint32_t)L.P * 16u);             // float4 [P] centroid x, y, mass bits, n_cells bits
  o.pcell = p;   p += ag_align16((uint32_t)L.P * 16u);             // float4 [P] cell of a lane-ticked player: x, y, mass bits, valid (>= 0)
  return p;
}

#ifdef __CUDACC__
struct WarpSmem {
  uint8_t* base;
  const SmemOff* o;
#define AG_SM(name, type) __device__ __forceinline__ type* name() const { return reinterpret_cast<type*>(base + o->name); }
  AG_SM(hcnt, uint16_t) AG_SM(htmp, uint32_t) AG_SM(hsorted, uint16_t) AG_SM(spel, float2) AG_SM(mbar, uint64_t)
  AG_SM(cellref, uint16_t) AG_SM(rows, int16_t) AG_SM(strip, uint16_t) AG_SM(hitq, uint16_t) AG_SM(snap, float4)
  AG_SM(vcache, float4) AG_SM(psum, float4) AG_SM(pcell, float4)
  AG_SM(pairs, uint4) AG_SM(reskeys, uint16_t) AG_SM(resorder, uint16_t)
  AG_SM(cand, uint2) AG_SM(prem, uint16_t) AG_SM(vrem, uint16_t) AG_SM(lprem, uint16_t)
#undef AG_SM
};
#endif

}  // namespace ag
