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

static_assert(kPairCap * 16 + 2 * ((kPairCap * 2 + 15) / 16 * 16) <= kCandCap * 8 + (kPremCap * 2 + 15) / 16 * 16 + 32 * kLaneCand * 2,
              "exact-sweep scratch must fit in the cand/prem/lprem region it aliases");

// Fills the offsets; returns the bytes one warp needs.
inline uint32_t make_smem_offsets(const agarcl_layout& L, int HG, SmemOff& o) {
  uint32_t p = 0;
  // pellet spatial hash: persists across the ticks of a launch, patched on removals
  o.hcnt = p;    p += ag_align16((uint32_t)(HG * HG) * 4u);        // u32 [HG*HG]   counts -> offsets -> cell ends
  o.hsorted = p; p += ag_align16((uint32_t)L.cap_pellets * 2u);    // u16 [cap_pellets] pellet indices grouped by hash cell
  // the instance's pellet array itself (index order), brought in by ONE TMA bulk load when the warp takes the
  // instance and written back by one bulk store if a pellet was eaten or spawned: every pellet access of the
  // ticks and of the observation scatter is a shared-memory access, not a trip to L2 / HBM behind the obs stores
  o.spel = p;    p += ag_align16((uint32_t)L.cap_pellets * 8u);    // float2 [cap_pellets]
  o.mbar = p;    p += 16u;                                         // u64 mbarrier of the bulk load
  // players_collision
  o.cellref = p; p += ag_align16(kCellRefCap * 2u);                // u16 [kCellRefCap] (player << 8 | cell) in snapshot order
  o.rows = p;    p += ag_align16(kCellRefCap * 2u);                // i16 [kCellRefCap] strip id
  o.strip = p;   p += ag_align16(kCellRefCap * 2u);                // u16 [kCellRefCap] one strip sorted by y
  o.hitq = p;    p += ag_align16(kPairCap * 2u);                   // u16 [kPairCap] queries flagged by the pre-test
  o.snap = p;    p += ag_align16(kSnapCap * 16u);                  // float4 [kSnapCap] snapshot: x, y, mass bits, player
  // live across ticks
  o.vcache = p;  p += ag_align16((uint32_t)L.cap_viruses * 16u);   // float4 [cap_viruses] x, y, radius, mass bits
  o.psum = p;    p += ag_align16((uint32_t)L.P * 16u);             // float4 [P] centroid x, y, mass bits, n_cells bits
  o.pcell = p;   p += ag_align16((uint32_t)L.P * 16u);             // float4 [P] cell of a lane-ticked player: x, y, mass bits, valid (>= 0)
  // player loop scratch; the exact collision sweep (pairs, reskeys, resorder) aliases it afterwards
  o.pairs = p;                                                     // uint4 [kPairCap] PairRec
  o.reskeys = o.pairs + kPairCap * 16u;                            // u16 [kPairCap]
  o.resorder = o.reskeys + ag_align16(kPairCap * 2u);              // u16 [kPairCap]
  o.cand = p;    p += ag_align16(kCandCap * 8u);                   // uint2 [kCandCap] (order key, d^2 bits)
  o.prem = p;    p += ag_align16(kPremCap * 2u);                   // u16 [kPremCap] pellets_to_remove
  o.vrem = p;    p += ag_align16(kVremCap * 2u);                   // u16 [kVremCap] viruses_to_remove
  o.lprem = p;   p += ag_align16(32u * kLaneCand * 2u);            // u16 [32][kLaneCand] pellets eaten by each lane's player
  return p;
}

#ifdef __CUDACC__
struct WarpSmem {
  uint8_t* base;
  const SmemOff* o;
#define AG_SM(name, type) __device__ __forceinline__ type* name() const { return reinterpret_cast<type*>(base + o->name); }
  AG_SM(hcnt, uint32_t) AG_SM(hsorted, uint16_t) AG_SM(spel, float2) AG_SM(mbar, uint64_t)
  AG_SM(cellref, uint16_t) AG_SM(rows, int16_t) AG_SM(strip, uint16_t) AG_SM(hitq, uint16_t) AG_SM(snap, float4)
  AG_SM(vcache, float4) AG_SM(psum, float4) AG_SM(pcell, float4)
  AG_SM(pairs, uint4) AG_SM(reskeys, uint16_t) AG_SM(resorder, uint16_t)
  AG_SM(cand, uint2) AG_SM(prem, uint16_t) AG_SM(vrem, uint16_t) AG_SM(lprem, uint16_t)
#undef AG_SM
};
#endif

}  // namespace ag
