// mirror.cu — host-resident observation mirror: the reference-facing hand-off of GridObservation frames to
// HOST memory (get_state, environment/bindings.cpp:67-91, copies 512 KB per agent per step) without moving
// the dense tensor over PCIe.
//
// A grid observation is almost entirely structure: channel 0 of every frame is the out-of-bounds mask, which
// is the union of whole rows and whole columns (GridObservation::_mark_out_of_bounds,
// environment/envs/GridEnvironment.hpp:235-248: a grid point is marked iff its x or its y leaves the arena),
// and the other channels hold a few hundred non-zero elements (one per in-view entity,
// _store_entities :212-232).  So after every step
//   1. k_pack (one CTA per agent image) reads the dense device observation once and emits, per image,
//      the row/column bit masks of every mask channel and the (offset, value) list of all other non-zeros,
//      packed into one device array through a single atomic cursor;
//   2. the lists (a few MB instead of 2.1 GB at configs[1]) are copied to pinned host memory;
//   3. a pool of host threads brings the library-owned dense mirror [N*A, frames*C, G, G] up to date in place:
//      zero the previous step's entries, rewrite only the mask rows/columns that changed, store the new entries.
// An image that does not fit the scheme (entry capacity exceeded, or a mask channel that is not a row/column
// union) is copied densely instead, so the mirror is ALWAYS identical to the device tensor — the parity test
// tests/test_gpu_mirror.py compares them element for element after every step.
#include "mirror.h"

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "host_util.h"

namespace ag {

constexpr int kPackThreads = 256;
constexpr uint32_t kDense = 0xFFFFFFFFu;  // img_count marker: this image takes the dense copy

struct PackParams {
  const void* obs;
  int32_t n_img, CH, C, G, MW;  // MW: 32-bit words per row (or column) mask
  uint32_t cap_img, cap_total;
  uint2* entries;               // [cap_total] (offset inside the image, value)
  uint32_t* img_count;          // [n_img] entries of the image, or kDense
  uint32_t* img_base;           // [n_img] first entry of the image
  uint32_t* total;              // atomic cursor into entries
  uint32_t* masks;              // [n_img][frames][2][MW]: bit i of the row mask = row i is out of bounds
};

template <typename T> struct Vec;
template <> struct Vec<int32_t> {
  static constexpr int kN = 4;
  static __device__ __forceinline__ void unpack(const int4& q, int (&e)[4]) { e[0] = q.x; e[1] = q.y; e[2] = q.z; e[3] = q.w; }
};
template <> struct Vec<int16_t> {
  static constexpr int kN = 8;
  static __device__ __forceinline__ void unpack(const int4& q, int (&e)[8]) {
    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      e[2 * k] = (int)(int16_t)(w[k] & 0xFFFF);
      e[2 * k + 1] = w[k] >> 16;
    }
  }
};

// One CTA per agent image.  HBM-bound: reads every element of the dense observation exactly once from DRAM
// (the mask channels a second time from L2); algorithmic bytes = sizeof(obs) in, lists out.
template <typename T>
__global__ void __launch_bounds__(kPackThreads) k_pack(const PackParams P) {
  constexpr int VE = Vec<T>::kN;
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* s_ent = reinterpret_cast<uint2*>(smem);
  uint8_t* row_any = smem + (size_t)P.cap_img * sizeof(uint2);  // row i holds an in-bounds (zero) element
  uint8_t* col_any = row_any + P.G;
  __shared__ uint32_t s_cnt, s_bad, s_base;
  const uint32_t tid = threadIdx.x, img = blockIdx.x;
  const uint32_t G = (uint32_t)P.G, plane = G * G, pv = plane / VE;
  const T* base = reinterpret_cast<const T*>(P.obs) + (size_t)img * P.CH * plane;
  if (tid == 0) { s_cnt = 0; s_bad = 0; }
  for (uint32_t i = tid; i < 2 * G; i += kPackThreads) row_any[i] = 0;
  __syncthreads();
  const int frames = P.CH / P.C;
  for (int f = 0; f < frames; f++) {
    const int4* p = reinterpret_cast<const int4*>(base + (size_t)f * P.C * plane);
    for (uint32_t v = tid; v < pv; v += kPackThreads) {
      int e[VE];
      Vec<T>::unpack(__ldg(p + v), e);
#pragma unroll
      for (int k = 0; k < VE; k++)
        if (e[k] == 0) {
          const uint32_t idx = v * VE + k;
          row_any[idx / G] = 1;
          col_any[idx % G] = 1;
        }
    }
    __syncthreads();
    bool bad = false;
    for (uint32_t v = tid; v < pv; v += kPackThreads) {
      int e[VE];
      Vec<T>::unpack(__ldg(p + v), e);
#pragma unroll
      for (int k = 0; k < VE; k++) {
        const uint32_t idx = v * VE + k;
        const int expect = (row_any[idx / G] & col_any[idx % G]) ? 0 : -1;
        bad |= (e[k] != expect);
      }
    }
    if (bad) s_bad = 1;
    uint32_t* mk = P.masks + ((size_t)img * frames + f) * 2 * P.MW;
    for (uint32_t w = tid; w < 2u * P.MW; w += kPackThreads) {
      const uint8_t* any = w < (uint32_t)P.MW ? row_any : col_any;
      const uint32_t w0 = (w % P.MW) * 32;
      uint32_t bits = 0;
      for (uint32_t b = 0; b < 32 && w0 + b < G; b++)
        if (!any[w0 + b]) bits |= 1u << b;
      mk[w] = bits;
    }
    __syncthreads();
    for (uint32_t i = tid; i < 2 * G; i += kPackThreads) row_any[i] = 0;
    __syncthreads();
  }
  // every other channel: non-zero elements are rare, four 16-byte loads in flight per thread
  for (int c = 0; c < P.CH; c++) {
    if (c % P.C == 0) continue;
    const int4* p = reinterpret_cast<const int4*>(base + (size_t)c * plane);
    for (uint32_t v0 = tid; v0 < pv; v0 += 4 * kPackThreads) {
      int4 q[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t v = v0 + u * kPackThreads;
        q[u] = v < pv ? __ldcs(p + v) : make_int4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if ((q[u].x | q[u].y | q[u].z | q[u].w) == 0) continue;
        int e[VE];
        Vec<T>::unpack(q[u], e);
#pragma unroll
        for (int k = 0; k < VE; k++)
          if (e[k] != 0) {
            const uint32_t slot = atomicAdd(&s_cnt, 1u);
            if (slot < P.cap_img) s_ent[slot] = make_uint2((uint32_t)c * plane + (v0 + u * kPackThreads) * VE + k, (uint32_t)e[k]);
          }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t cnt = s_cnt, b0 = 0;
    if (s_bad || cnt > P.cap_img) cnt = kDense;
    else if (cnt) {
      b0 = atomicAdd(P.total, cnt);
      if (b0 + cnt > P.cap_total) cnt = kDense;
    }
    P.img_count[img] = cnt;
    P.img_base[img] = b0;
    s_cnt = cnt;
    s_base = b0;
  }
  __syncthreads();
  const uint32_t cnt = s_cnt;
  if (cnt == kDense) return;
  for (uint32_t k = tid; k < cnt; k += kPackThreads) P.entries[s_base + k] = s_ent[k];
}

// ------------------------------------------------------------------------------------------------ host pool
// Every worker and the caller run the same job, which pulls chunks from an atomic counter.
class Pool {
 public:
  explicit Pool(int workers) {
    for (int i = 0; i < workers; i++) th_.emplace_back([this] { loop(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_start_.notify_all();
    for (auto& t : th_) t.join();
  }
  int threads() const { return (int)th_.size() + 1; }
  void run(const std::function<void()>& job) {
    {
      std::lock_guard<std::mutex> g(mu_);
      job_ = &job;
      pending_ = (int)th_.size();
      gen_++;
    }
    cv_start_.notify_all();
    job();
    std::unique_lock<std::mutex> g(mu_);
    cv_done_.wait(g, [this] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void()>* job;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_start_.wait(g, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        job = job_;
      }
      (*job)();
      {
        std::lock_guard<std::mutex> g(mu_);
        if (--pending_ == 0) cv_done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_start_, cv_done_;
  const std::function<void()>* job_ = nullptr;
  uint64_t gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};

struct HostMirror {
  int n_img = 0, CH = 0, C = 0, G = 0, frames = 0, MW = 0, dtype = 0;
  size_t esz = 4, img_elems = 0, img_bytes = 0;
  uint32_t cap_img = 0, cap_total = 0;
  void* h_obs = nullptr;          // pinned [n_img][CH][G][G]
  uint2* d_entries = nullptr;
  uint32_t* d_meta = nullptr;     // [count n_img][base n_img][total, pad x3][masks n_img*frames*2*MW]
  size_t meta_words = 0;
  uint2* h_entries[2] = {nullptr, nullptr};  // pinned staging; [cur] = this step's lists, [cur^1] = the previous step's
  uint32_t* h_meta[2] = {nullptr, nullptr};
  int cur = 0;
  size_t guess = 0;               // entries fetched together with the counts (the remainder, if any, in a second copy)
  Pool* pool = nullptr;
  MirrorStats stats{};
  uint32_t* count(int w) const { return h_meta[w]; }
  uint32_t* base(int w) const { return h_meta[w] + n_img; }
  uint32_t* total(int w) const { return h_meta[w] + 2 * (size_t)n_img; }
  uint32_t* masks(int w) const { return h_meta[w] + 2 * (size_t)n_img + 4; }
};

static inline bool bit(const uint32_t* m, int i) { return (m[i >> 5] >> (i & 31)) & 1u; }

// Channel 0 of one frame: element (i, j) is -1 iff row i or column j is out of bounds.
template <typename T>
static void apply_mask_delta(T* p, const uint32_t* om, const uint32_t* nm, int G, int MW) {
  const uint32_t *orow = om, *ocol = om + MW, *nrow = nm, *ncol = nm + MW;
  const bool col_changed = std::memcmp(ocol, ncol, (size_t)MW * 4) != 0;
  for (int i = 0; i < G; i++) {
    const bool ro = bit(orow, i), rn = bit(nrow, i);
    T* r = p + (size_t)i * G;
    if (rn) {
      if (!ro) std::memset(r, 0xFF, (size_t)G * sizeof(T));
      continue;
    }
    if (ro) {
      for (int j = 0; j < G; j++) r[j] = bit(ncol, j) ? (T)-1 : (T)0;
      continue;
    }
    if (col_changed)
      for (int w = 0; w < MW; w++)
        for (uint32_t d = ocol[w] ^ ncol[w]; d; d &= d - 1) {
          const int j = w * 32 + __builtin_ctz(d);
          r[j] = bit(ncol, j) ? (T)-1 : (T)0;
        }
  }
}

template <typename T>
static void expand_range(HostMirror* m, int lo, int hi, const uint32_t* zero_masks) {
  const int cur = m->cur, prev = cur ^ 1;
  const size_t mw2 = 2 * (size_t)m->MW, plane = (size_t)m->G * m->G;
  for (int img = lo; img < hi; img++) {
    const uint32_t cnt = m->count(cur)[img];
    if (cnt == kDense) continue;  // the dense copy of this image is already in flight
    T* p = reinterpret_cast<T*>(m->h_obs) + (size_t)img * m->img_elems;
    const uint32_t pcnt = m->count(prev)[img];
    const uint32_t* om = m->masks(prev) + (size_t)img * m->frames * mw2;
    const uint32_t* nm = m->masks(cur) + (size_t)img * m->frames * mw2;
    if (pcnt == kDense) {
      std::memset(p, 0, m->img_bytes);
      om = nullptr;
    } else {
      const uint2* pe = m->h_entries[prev] + m->base(prev)[img];
      for (uint32_t k = 0; k < pcnt; k++) p[pe[k].x] = 0;
    }
    for (int f = 0; f < m->frames; f++) {
      const uint32_t* o = om ? om + f * mw2 : zero_masks;
      if (std::memcmp(o, nm + f * mw2, mw2 * 4) != 0) apply_mask_delta<T>(p + (size_t)f * m->C * plane, o, nm + f * mw2, m->G, m->MW);
    }
    const uint2* ne = m->h_entries[cur] + m->base(cur)[img];
    for (uint32_t k = 0; k < cnt; k++) p[ne[k].x] = (T)(int32_t)ne[k].y;
  }
}

void mirror_destroy(HostMirror* m) {
  if (!m) return;
  delete m->pool;
  cudaFree(m->d_entries);
  cudaFree(m->d_meta);
  for (int w = 0; w < 2; w++) {
    cudaFreeHost(m->h_entries[w]);
    cudaFreeHost(m->h_meta[w]);
  }
  cudaFreeHost(m->h_obs);
  delete m;
}

HostMirror* mirror_create(int n_img, int CH, int C, int G, int dtype) {
  const size_t esz = dtype == AGARCL_OBS_I16 ? 2 : 4;
  if (((size_t)G * G * esz) % 16 != 0 || G > 4096) {
    agarcl_set_error(AGARCL_ERR_INVALID, "the host mirror needs grid planes that are multiples of 16 bytes (grid_size %d)", G);
    return nullptr;
  }
  HostMirror* m = new (std::nothrow) HostMirror();
  if (!m) {
    agarcl_set_error(AGARCL_ERR_NOMEM, "out of host memory");
    return nullptr;
  }
  m->n_img = n_img; m->CH = CH; m->C = C; m->G = G; m->frames = CH / C; m->MW = (G + 31) / 32; m->dtype = dtype;
  m->esz = esz;
  m->img_elems = (size_t)CH * G * G;
  m->img_bytes = m->img_elems * esz;
  m->cap_img = 2048;
  if (const char* e = std::getenv("AGARCL_MIRROR_CAP_IMG")) m->cap_img = (uint32_t)std::atoi(e);
  uint64_t ct = (uint64_t)n_img * (m->cap_img < 1024 ? m->cap_img : 1024);
  if (ct < 4096) ct = 4096;
  m->cap_total = (uint32_t)(ct > 0x7FFFFFFFull ? 0x7FFFFFFFull : ct);
  m->meta_words = 2 * (size_t)n_img + 4 + (size_t)n_img * m->frames * 2 * m->MW;
  bool ok = cudaHostAlloc(&m->h_obs, (size_t)n_img * m->img_bytes, cudaHostAllocDefault) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&m->d_entries, (size_t)m->cap_total * sizeof(uint2)) == cudaSuccess;
  ok = ok && cudaMalloc((void**)&m->d_meta, m->meta_words * 4) == cudaSuccess;
  for (int w = 0; w < 2 && ok; w++) {
    ok = ok && cudaHostAlloc((void**)&m->h_entries[w], (size_t)m->cap_total * sizeof(uint2), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaHostAlloc((void**)&m->h_meta[w], m->meta_words * 4, cudaHostAllocDefault) == cudaSuccess;
    if (ok) std::memset(m->h_meta[w], 0, m->meta_words * 4);  // no entries, nothing out of bounds == the all-zero mirror
  }
  if (!ok) {
    agarcl_set_error(AGARCL_ERR_NOMEM, "host mirror allocation failed (%zu B pinned): %s", (size_t)n_img * m->img_bytes,
                     cudaGetErrorString(cudaGetLastError()));
    mirror_destroy(m);
    return nullptr;
  }
  int nt = (int)std::thread::hardware_concurrency();
  if (const char* e = std::getenv("AGARCL_HOST_THREADS")) nt = std::atoi(e);
  nt = nt < 1 ? 1 : (nt > 64 ? 64 : nt);
  if (nt > n_img) nt = n_img;
  m->pool = new Pool(nt - 1);
  m->stats.host_threads = (uint64_t)nt;
  m->guess = (size_t)n_img * 64;
  // first touch of the mirror by the threads that will write it
  {
    std::atomic<int> next{0};
    m->pool->run([&] {
      for (int i; (i = next.fetch_add(8)) < n_img;) {
        const int hi = i + 8 < n_img ? i + 8 : n_img;
        std::memset((uint8_t*)m->h_obs + (size_t)i * m->img_bytes, 0, (size_t)(hi - i) * m->img_bytes);
      }
    });
  }
  return m;
}

void* mirror_ptr(HostMirror* m) { return m->h_obs; }
void mirror_stats(const HostMirror* m, MirrorStats* out) { *out = m->stats; }

#define MCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return agarcl_set_error(AGARCL_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));   \
  } while (0)

int mirror_sync(HostMirror* m, const void* d_obs, cudaStream_t s) {
  m->cur ^= 1;
  const int cur = m->cur;
  PackParams P;
  P.obs = d_obs;
  P.n_img = m->n_img; P.CH = m->CH; P.C = m->C; P.G = m->G; P.MW = m->MW;
  P.cap_img = m->cap_img; P.cap_total = m->cap_total;
  P.entries = m->d_entries;
  P.img_count = m->d_meta;
  P.img_base = m->d_meta + m->n_img;
  P.total = m->d_meta + 2 * (size_t)m->n_img;
  P.masks = m->d_meta + 2 * (size_t)m->n_img + 4;
  MCK(cudaMemsetAsync(P.total, 0, 16, s));
  const size_t smem = (size_t)m->cap_img * sizeof(uint2) + 2 * (size_t)m->G + 16;
  if (m->dtype == AGARCL_OBS_I16) {
    if (smem > 48 * 1024) MCK(cudaFuncSetAttribute(k_pack<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pack<int16_t><<<m->n_img, kPackThreads, smem, s>>>(P);
  } else {
    if (smem > 48 * 1024) MCK(cudaFuncSetAttribute(k_pack<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pack<int32_t><<<m->n_img, kPackThreads, smem, s>>>(P);
  }
  MCK(cudaGetLastError());
  MCK(cudaMemcpyAsync(m->h_meta[cur], m->d_meta, m->meta_words * 4, cudaMemcpyDeviceToHost, s));
  size_t got = m->guess < m->cap_total ? m->guess : m->cap_total;
  MCK(cudaMemcpyAsync(m->h_entries[cur], m->d_entries, got * sizeof(uint2), cudaMemcpyDeviceToHost, s));
  MCK(cudaStreamSynchronize(s));
  uint64_t d2h = m->meta_words * 4 + got * sizeof(uint2);
  size_t total = *m->total(cur);
  if (total > m->cap_total) total = m->cap_total;
  if (total > got) {
    MCK(cudaMemcpyAsync(m->h_entries[cur] + got, m->d_entries + got, (total - got) * sizeof(uint2), cudaMemcpyDeviceToHost, s));
    d2h += (total - got) * sizeof(uint2);
    MCK(cudaStreamSynchronize(s));
  }
  m->guess = total + total / 8 + 4096;
  // dense copies first (asynchronous into the pinned mirror), they overlap the list expansion below
  uint64_t dense = 0;
  const uint32_t* cnt = m->count(cur);
  for (int i = 0; i < m->n_img; i++)
    if (cnt[i] == kDense) {
      MCK(cudaMemcpyAsync((uint8_t*)m->h_obs + (size_t)i * m->img_bytes, (const uint8_t*)d_obs + (size_t)i * m->img_bytes,
                          m->img_bytes, cudaMemcpyDeviceToHost, s));
      dense++;
    }
  std::vector<uint32_t> zero_masks(2 * (size_t)m->MW, 0u);
  std::atomic<int> next{0};
  const int chunk = 8;
  m->pool->run([&] {
    for (int i; (i = next.fetch_add(chunk)) < m->n_img;) {
      const int hi = i + chunk < m->n_img ? i + chunk : m->n_img;
      if (m->dtype == AGARCL_OBS_I16) expand_range<int16_t>(m, i, hi, zero_masks.data());
      else expand_range<int32_t>(m, i, hi, zero_masks.data());
    }
  });
  if (dense) MCK(cudaStreamSynchronize(s));
  m->stats.entries = total;
  m->stats.dense_images = dense;
  m->stats.d2h_bytes = d2h + dense * m->img_bytes;
  return AGARCL_OK;
}

}  // namespace ag
