// sim_shared.cuh — shared-memory carve-up of one warp (= one game instance) in k_step.
#pragma once
#include <cstdint>

#include "../../include/agarcl_b200.h"
#include "sim_params.h"

namespace ag {

struct WarpSmem {
  // pellet spatial hash, rebuilt every tick (valid during the player loop)
  uint32_t* hcnt;      // [HG*HG]   counts -> offsets -> cell ends
  uint16_t* hsorted;   // [cap_pellets] pellet indices grouped by hash cell
  uint32_t* hq;        // [cap_pellets] same order: pellet position quantised to 2 x 16 bits (x | y << 16)
  // (the hash persists across the ticks of a launch and is patched on removals, so nothing aliases it)
  // players_collision scratch
  uint16_t* cellref;   // [kCellRefCap]
  int16_t* rows;       // [kCellRefCap]
  uint16_t* strip;     // [kCellRefCap]
  uint4* pairs;        // [kPairCap] PairRec          } exact sweep only: these three alias cand/prem/lprem,
  uint16_t* reskeys;   // [kPairCap]                  } which are dead once the removals are applied
  uint16_t* resorder;  // [kPairCap]                  }
  uint16_t* hitq;      // [kPairCap]
  // live for the whole launch
  float4* vcache;      // [cap_viruses] x, y, radius, mass bits
  float4* psum;        // [P] centroid x, y, mass bits, n_cells bits
  float4* pcell;       // [P] the cell of a lane-ticked single-cell player: x, y, mass bits, valid (>= 0)
  uint2* cand;         // [kCandCap] (order key, d^2 bits)
  uint16_t* prem;      // [kPremCap]
  uint16_t* vrem;      // [kVremCap]
  uint16_t* lprem;     // [32][kLaneCand] pellets eaten by each lane's player in the lane-per-player phase
  float4* snap;        // [kSnapCap] players_collision snapshot: x, y, mass bits, player
};

__host__ __device__ inline uint32_t ag_align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline uint32_t hash_region_bytes(const agarcl_layout& L, int HG) {
  return ag_align16((uint32_t)(HG * HG) * 4u) + ag_align16((uint32_t)L.cap_pellets * 2u) + ag_align16((uint32_t)L.cap_pellets * 4u);
}
__host__ __device__ inline uint32_t coll_region_bytes() {
  return ag_align16(kCellRefCap * 2u) * 3u + ag_align16(kPairCap * 2u);
}
static_assert(kPairCap * 16 + 2 * ((kPairCap * 2 + 15) / 16 * 16) <= kCandCap * 8 + (kPremCap * 2 + 15) / 16 * 16 + 32 * kLaneCand * 2,
              "exact-sweep scratch must fit in the cand/prem/lprem region it aliases");
__host__ __device__ inline uint32_t warp_smem_bytes(const agarcl_layout& L, int HG) {
  return hash_region_bytes(L, HG) + coll_region_bytes() + ag_align16((uint32_t)L.cap_viruses * 16u) +
         2u * ag_align16((uint32_t)L.P * 16u) + ag_align16(kCandCap * 8u) + ag_align16(kPremCap * 2u) + ag_align16(kVremCap * 2u) +
         ag_align16(32u * kLaneCand * 2u) + ag_align16(kSnapCap * 16u);
}

__device__ inline WarpSmem carve_warp_smem(uint8_t* base, const agarcl_layout& L, int HG) {
  WarpSmem s;
  uint8_t* p = base;
  s.hcnt = reinterpret_cast<uint32_t*>(p);
  s.hsorted = reinterpret_cast<uint16_t*>(p + ag_align16((uint32_t)(HG * HG) * 4u));
  s.hq = reinterpret_cast<uint32_t*>(p + ag_align16((uint32_t)(HG * HG) * 4u) + ag_align16((uint32_t)L.cap_pellets * 2u));
  uint8_t* q = base + hash_region_bytes(L, HG);
  s.cellref = reinterpret_cast<uint16_t*>(q); q += ag_align16(kCellRefCap * 2u);
  s.rows = reinterpret_cast<int16_t*>(q);     q += ag_align16(kCellRefCap * 2u);
  s.strip = reinterpret_cast<uint16_t*>(q);   q += ag_align16(kCellRefCap * 2u);
  s.hitq = reinterpret_cast<uint16_t*>(q);
  p += hash_region_bytes(L, HG) + coll_region_bytes();
  s.vcache = reinterpret_cast<float4*>(p); p += ag_align16((uint32_t)L.cap_viruses * 16u);
  s.psum = reinterpret_cast<float4*>(p);   p += ag_align16((uint32_t)L.P * 16u);
  s.pcell = reinterpret_cast<float4*>(p);  p += ag_align16((uint32_t)L.P * 16u);
  {  // exact-sweep scratch over cand | prem | vrem | lprem (contiguous)
    uint8_t* a = p;
    s.pairs = reinterpret_cast<uint4*>(a);       a += kPairCap * 16u;
    s.reskeys = reinterpret_cast<uint16_t*>(a);  a += ag_align16(kPairCap * 2u);
    s.resorder = reinterpret_cast<uint16_t*>(a);
  }
  s.cand = reinterpret_cast<uint2*>(p);    p += ag_align16(kCandCap * 8u);
  s.prem = reinterpret_cast<uint16_t*>(p); p += ag_align16(kPremCap * 2u);
  s.vrem = reinterpret_cast<uint16_t*>(p);  p += ag_align16(kVremCap * 2u);
  s.lprem = reinterpret_cast<uint16_t*>(p); p += ag_align16(32u * kLaneCand * 2u);
  s.snap = reinterpret_cast<float4*>(p);
  return s;
}

}  // namespace ag
