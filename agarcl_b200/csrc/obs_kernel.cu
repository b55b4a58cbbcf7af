// obs_kernel.cu — GridObservation::add_frame for every agent of every instance, for sm_100a.
//
// Reference: environment/envs/GridEnvironment.hpp — add_frame 91-123, _store_entities 212-232,
// _mark_out_of_bounds 235-248, _view_size 251-254, _world_to_grid 257-267, _grid_to_world 270-279,
// _index 282-287, _inside_grid 290-292, _in_bounds 294-296.
//
// One CTA renders one agent's [C][G][G] frame (512 KiB for the default 8 x 128 x 128 int32) and is
// purely HBM-write bound, so the layout of the work is: (1) every output element is produced by a
// 16-byte streaming vector store exactly once — channel 0 (out-of-bounds mask) is separable into a
// row predicate and a column predicate held in shared memory, all other channels are zero — then
// (2) after one block barrier the few hundred in-view entities are scattered with L2 atomics
// (sum / max / encoded min / highest-index-wins), which touches lines that are still in L2.
// Output layout is the reference's: index = c*G*G + gx*G + gy  (x is the slow axis).
#include <cuda_runtime.h>

#include "sim_params.h"

namespace ag {

constexpr int kObsThreads = 256;


template <typename T> struct ObsOps;
template <> struct ObsOps<int32_t> {
  static __device__ __forceinline__ void store(int32_t* p, int v) { *p = v; }
  static __device__ __forceinline__ void add(int32_t* p, int v) { atomicAdd(p, v); }
  static __device__ __forceinline__ void maxs(int32_t* p, int v) { atomicMax(p, v); }
  // min over non-zero entries, phase A: keep max of the complement (cells start at 0)
  static __device__ __forceinline__ void min_enc(int32_t* p, int v) { atomicMax(reinterpret_cast<unsigned*>(p), ~(unsigned)v); }
  static __device__ __forceinline__ int min_dec(int32_t e) { return (int)~(unsigned)e; }
};
// opt-in narrow dtype: 16-bit cells updated through 32-bit CAS on the containing word, saturating
template <> struct ObsOps<int16_t> {
  template <typename F> static __device__ __forceinline__ void rmw(int16_t* p, F f) {
    unsigned* w = reinterpret_cast<unsigned*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const int sh = (reinterpret_cast<uintptr_t>(p) & 2) ? 16 : 0;
    unsigned old = *w, assumed;
    do {
      assumed = old;
      int16_t cur = (int16_t)((assumed >> sh) & 0xffffu);
      int16_t nv = f(cur);
      unsigned repl = (assumed & ~(0xffffu << sh)) | (((unsigned)(uint16_t)nv) << sh);
      old = atomicCAS(w, assumed, repl);
    } while (old != assumed);
  }
  static __device__ __forceinline__ int16_t sat(int v) { return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
  static __device__ __forceinline__ void store(int16_t* p, int v) { int16_t s = sat(v); rmw(p, [s](int16_t) { return s; }); }
  static __device__ __forceinline__ void add(int16_t* p, int v) { rmw(p, [v](int16_t c) { return sat((int)c + v); }); }
  static __device__ __forceinline__ void maxs(int16_t* p, int v) { int16_t s = sat(v); rmw(p, [s](int16_t c) { return c > s ? c : s; }); }
  static __device__ __forceinline__ void min_enc(int16_t* p, int v) {
    int16_t s = (int16_t)~sat(v);  // complement is negative: smaller mass -> larger complement
    rmw(p, [s](int16_t c) { return (c == 0 || s > c) ? s : c; });
  }
  static __device__ __forceinline__ int min_dec(int16_t e) { return (int)(int16_t)~e; }
};

template <typename T>
__global__ void __launch_bounds__(kObsThreads) k_obs(const __grid_constant__ ObsParams P) {
  __shared__ float s_px, s_py, s_view;
  __shared__ int s_alive;
  __shared__ __align__(16) T s_ymask[1024];  // 0 / -1 per gy
  __shared__ T s_xmask[1024];                // 0 / -1 per gx
  const int A = P.L.A, G = P.G, C = P.C;
  const int inst = blockIdx.x / A, agent = blockIdx.x % A;
  if (P.mask && !P.mask[inst]) return;
  const uint8_t* blob = P.state + (size_t)inst * P.L.stride;
  const agarcl_inst_hdr* hdr = reinterpret_cast<const agarcl_inst_hdr*>(blob + P.L.off_hdr);
  const agarcl_player* players = reinterpret_cast<const agarcl_player*>(blob + P.L.off_players);
  const agarcl_cell* cells = reinterpret_cast<const agarcl_cell*>(blob + P.L.off_cells);
  T* out = reinterpret_cast<T*>(P.obs) + ((size_t)blockIdx.x * P.frames + P.frame) * (size_t)C * G * G;
  const int tid = threadIdx.x;
  // players respawned by the step that just ran are still dead in that step's observation
  const unsigned long long resp = P.pre_respawn ? ((unsigned long long)hdr->respawned_hi << 32 | hdr->respawned_lo) : 0ull;
  auto ncells = [&](int p) -> int { return ((resp >> p) & 1ull) ? 0 : players[p].n_cells; };

  if (tid == 0) {
    // Player::x/y/mass in cell order (Player.hpp:102-126); _view_size = clamp(2*mass, 100, 300)
    const agarcl_cell* pc = cells + (size_t)agent * AGARCL_MAX_CELLS;
    int n = ncells(agent);
    float xs = 0.f, ys = 0.f;
    uint32_t tot = 0;
    for (int i = 0; i < n; i++) { xs += pc[i].x * (float)pc[i].mass; ys += pc[i].y * (float)pc[i].mass; tot += pc[i].mass; }
    float fm = (float)tot;
    s_px = xs / fm;
    s_py = ys / fm;
    s_view = clamp_std((float)(2u * tot), 100.0f, 300.0f);
    s_alive = n > 0;
  }
  __syncthreads();
  const float px = s_px, py = s_py, view = s_view, W = P.W;
  const float centering = (float)(G / 2.0);
  // _grid_to_world + _in_bounds, separable in i (x) and j (y)
  for (int i = tid; i < G; i += kObsThreads) {
    float d = (float)i - centering;
    float wx = px + d * view / (float)G, wy = py + d * view / (float)G;
    s_xmask[i] = (0 <= wx && wx < W) ? (T)0 : (T)-1;
    s_ymask[i] = (0 <= wy && wy < W) ? (T)0 : (T)-1;
  }
  __syncthreads();

  // ---- phase 1: every element written once with 16-byte streaming stores
  constexpr int VE = 16 / (int)sizeof(T);  // elements per vector
  const size_t plane = (size_t)G * G;
  if (G % VE == 0) {
    const int vec_per_row = G / VE;
    const int nvec0 = (int)(plane / VE), nvec = C * nvec0;
    int4* out4 = reinterpret_cast<int4*>(out);
    for (int v = tid; v < nvec0; v += kObsThreads) {  // channel 0: out-of-bounds mask
      int i = v / vec_per_row, jv = v - i * vec_per_row;
      int4 ym = reinterpret_cast<const int4*>(s_ymask)[jv];
      int xm = (s_xmask[i] != 0) ? -1 : 0;
      __stcs(out4 + v, make_int4(ym.x | xm, ym.y | xm, ym.z | xm, ym.w | xm));
    }
    if (!P.skip_zero) {
      const int4 zero = make_int4(0, 0, 0, 0);
#pragma unroll 8
      for (int v = nvec0 + tid; v < nvec; v += kObsThreads) __stcs(out4 + v, zero);
    }
  } else {
    const size_t nel = (size_t)C * plane;
    for (size_t e = tid; e < nel; e += kObsThreads) {
      T val = 0;
      if (e < plane) { int i = (int)(e / G), j = (int)(e % G); val = (T)(s_xmask[i] | s_ymask[j]); }
      out[e] = val;
    }
  }
  if (!s_alive) return;  // dead agent: centroid is NaN, nothing lands inside the grid (quirk Q20)
  __syncthreads();

  // ---- phase 2: scatter entities
  auto grid_of = [&](float x, float y, int& gx, int& gy) -> bool {
    gx = to_int_x86((float)G * (x - px) / view + centering);
    gy = to_int_x86((float)G * (y - py) / view + centering);
    return 0 <= gx && gx < G && 0 <= gy && gy < G;
  };
  int channel = 0;
  if (P.observe_pellets) {
    const float2* pel = reinterpret_cast<const float2*>(blob + P.L.off_pellets);
    const int np = hdr->n_pellets;
    T* ch1 = out + (size_t)(channel + 1) * plane;
    T* ch2 = out + (size_t)(channel + 2) * plane;
    for (int k = tid; k < np; k += kObsThreads) {
      float2 q = pel[k];
      int gx, gy;
      if (grid_of(q.x, q.y, gx, gy)) {
        ObsOps<T>::store(ch1 + (size_t)gx * G + gy, 1);  // at_least_: data = mass (1)
        ObsOps<T>::add(ch2 + (size_t)gx * G + gy, 1);    // total_mass_
      }
    }
    channel += 2;
  }
  if (P.observe_viruses) {
    const agarcl_virus* vir = reinterpret_cast<const agarcl_virus*>(blob + P.L.off_viruses);
    const int nv = hdr->n_viruses;
    T* ch3 = out + (size_t)(channel + 1) * plane;
    T* ch4 = out + (size_t)(channel + 2) * plane;
    for (int k = tid; k < nv; k += kObsThreads) {
      int gx, gy;
      if (grid_of(vir[k].x, vir[k].y, gx, gy)) {
        ObsOps<T>::add(ch4 + (size_t)gx * G + gy, (int)vir[k].mass);
        // at_least_ keeps the LAST writer in index order: write only if no later virus shares the cell
        bool last = true;
        for (int k2 = k + 1; k2 < nv; k2++) {
          int hx, hy;
          if (grid_of(vir[k2].x, vir[k2].y, hx, hy) && hx == gx && hy == gy) { last = false; break; }
        }
        if (last) ObsOps<T>::store(ch3 + (size_t)gx * G + gy, (int)vir[k].mass);
      }
    }
    channel += 2;
  }
  if (P.observe_cells) {
    T* ch5 = out + (size_t)(channel + 1) * plane;
    const agarcl_cell* pc = cells + (size_t)agent * AGARCL_MAX_CELLS;
    const int n = ncells(agent);
    for (int k = tid; k < n; k += kObsThreads) {
      int gx, gy;
      if (grid_of(pc[k].x, pc[k].y, gx, gy)) ObsOps<T>::add(ch5 + (size_t)gx * G + gy, (int)pc[k].mass);
    }
    channel += 1;
  }
  if (P.observe_others) {
    T* ch6 = out + (size_t)(channel + 1) * plane;  // min over non-empty
    T* ch7 = out + (size_t)(channel + 2) * plane;  // max
    const int slots = P.L.P * AGARCL_MAX_CELLS;
    // phase A: encoded min + max
    for (int k = tid; k < slots; k += kObsThreads) {
      int p = k / AGARCL_MAX_CELLS, i = k % AGARCL_MAX_CELLS;
      if (p == agent || i >= ncells(p)) continue;
      const agarcl_cell* oc = cells + k;
      int gx, gy;
      if (grid_of(oc->x, oc->y, gx, gy)) {
        ObsOps<T>::min_enc(ch6 + (size_t)gx * G + gy, (int)oc->mass);
        ObsOps<T>::maxs(ch7 + (size_t)gx * G + gy, (int)oc->mass);
      }
    }
    __syncthreads();
    // phase B: read the encoded minimum; phase C: write it back decoded (all writers of a cell agree)
    constexpr int kMaxPer = (AGARCL_MAX_PLAYERS * AGARCL_MAX_CELLS + kObsThreads - 1) / kObsThreads;
    int dec[kMaxPer];
    int nd = 0;
    for (int k = tid; k < slots; k += kObsThreads, nd++) {
      dec[nd] = 0;
      int p = k / AGARCL_MAX_CELLS, i = k % AGARCL_MAX_CELLS;
      if (p == agent || i >= ncells(p)) continue;
      const agarcl_cell* oc = cells + k;
      int gx, gy;
      if (grid_of(oc->x, oc->y, gx, gy)) dec[nd] = ObsOps<T>::min_dec(__ldcg(ch6 + (size_t)gx * G + gy));
    }
    __syncthreads();
    nd = 0;
    for (int k = tid; k < slots; k += kObsThreads, nd++) {
      int p = k / AGARCL_MAX_CELLS, i = k % AGARCL_MAX_CELLS;
      if (p == agent || i >= ncells(p)) continue;
      const agarcl_cell* oc = cells + k;
      int gx, gy;
      if (grid_of(oc->x, oc->y, gx, gy)) ObsOps<T>::store(ch6 + (size_t)gx * G + gy, dec[nd]);
    }
  }
}

cudaError_t launch_obs(const ObsParams& P, cudaStream_t stream) {
  int ctas = P.N * P.L.A;
  if (P.obs_dtype == AGARCL_OBS_I16) k_obs<int16_t><<<ctas, kObsThreads, 0, stream>>>(P);
  else k_obs<int32_t><<<ctas, kObsThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace ag
