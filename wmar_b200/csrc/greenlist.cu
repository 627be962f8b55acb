// Greenlist engine: bit-exact replacement of GentimeWatermark._split_with_seed
// (wmar/watermarking/gentime_watermark.py:161-174) for the RANDOM and RANDOM_STRATIFIED splits.
//
// torch.Generator(cpu).manual_seed(s) seeds MT19937 with the low 32 bits of s; torch.randperm(n) on CPU is a forward
// Fisher-Yates  r[i] <-> r[i + rand32 % (n - i)],  i = 0..n-2.  Only the first int(n*gamma) outputs of each
// permutation are used, and position i is final after swap i, so each permutation is run for `need` swaps only while
// the generator is still advanced by the full n-1 draws (the dead permutation continues the alive one's stream).
//
// Output is a bitmask table [n_rows][ceil(V/32)]: row s is the greenlist of context sum s.
#include <thread>
#include <vector>

#include "common.cuh"

namespace wmar {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace wmar

using namespace wmar;

namespace {

// ------------------------------------------------------------------------------------------------ host build
struct MT {
    uint32_t s[624];
    int idx;
    void seed(uint32_t v) {
        s[0] = v;
        for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void twist() {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
            s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx = 0;
    }
    uint32_t next() {
        if (idx >= 624) twist();
        uint32_t y = s[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

// partial forward Fisher-Yates: first `need` entries final; consumes exactly max(n-1,0) draws
void partial_perm(MT &mt, int64_t n, int64_t need, std::vector<int32_t> &r) {
    r.resize((size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) r[i] = (int32_t)i;
    int64_t i = 0;
    for (; i + 1 < n && i < need; i++) {
        int64_t z = (int64_t)(mt.next() % (uint32_t)(n - i));
        int32_t t = r[i];
        r[i] = r[i + z];
        r[i + z] = t;
    }
    for (; i + 1 < n; i++) (void)mt.next();
}

void host_row(int64_t V, double gamma, int split, uint64_t seed, const int64_t *alive, int64_t n_alive,
              const int64_t *dead, int64_t n_dead, uint32_t *row, std::vector<int32_t> &pa, std::vector<int32_t> &pd) {
    const int64_t words = (V + 31) / 32;
    memset(row, 0, sizeof(uint32_t) * (size_t)words);
    MT mt;
    mt.seed((uint32_t)(seed & 0xffffffffu));
    const int64_t green = (int64_t)((double)V * gamma);
    if (split == WMAR_SPLIT_RANDOM) {
        int64_t need = green < V ? green : V;
        partial_perm(mt, V, need, pa);
        for (int64_t i = 0; i < need; i++) row[pa[i] >> 5] |= 1u << (pa[i] & 31);
        return;
    }
    int64_t n_ga = (int64_t)((double)n_alive * gamma);
    int64_t n_gd = green - n_ga;
    if (n_ga > n_alive) n_ga = n_alive;
    if (n_gd > n_dead) n_gd = n_dead;
    if (n_gd < 0) n_gd = 0;
    partial_perm(mt, n_alive, n_ga, pa);
    partial_perm(mt, n_dead, n_gd, pd);
    for (int64_t i = 0; i < n_ga; i++) {
        int64_t id = alive[pa[i]];
        if (id >= 0 && id < V) row[id >> 5] |= 1u << (id & 31);
    }
    for (int64_t i = 0; i < n_gd; i++) {
        int64_t id = dead[pd[i]];
        if (id >= 0 && id < V) row[id >> 5] |= 1u << (id & 31);
    }
}

// ------------------------------------------------------------------------------------------------ device build
// One CTA per table row.  MT19937 state in shared memory, regenerated 624 words at a time by the whole CTA (the twist
// is data-parallel in three phases: [0,227) reads only old words, [227,454) reads new [0,227), [454,624) reads new
// [227,397)); the swaps are inherently sequential and are done by thread 0 on a u16 permutation in shared memory.
constexpr int GL_THREADS = 256;

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far) {
    uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ void mt_regen(uint32_t *s, uint32_t *out) {
    const int tid = threadIdx.x;
    // phase A: i in [0,227): needs old s[i], s[i+1], s[i+397]
    uint32_t v[3];
    int cnt = 0;
    for (int i = tid; i < 227; i += GL_THREADS) v[cnt++] = mt_mix(s[i], s[i + 1], s[i + 397]);
    __syncthreads();
    cnt = 0;
    for (int i = tid; i < 227; i += GL_THREADS) s[i] = v[cnt++];
    __syncthreads();
    // phase B: i in [227,454): old s[i], old s[i+1], new s[i-227]
    cnt = 0;
    for (int i = 227 + tid; i < 454; i += GL_THREADS) v[cnt++] = mt_mix(s[i], s[i + 1], s[i - 227]);
    __syncthreads();
    cnt = 0;
    for (int i = 227 + tid; i < 454; i += GL_THREADS) s[i] = v[cnt++];
    __syncthreads();
    // phase C: i in [454,623): old s[i], old s[i+1], new s[i-227]; i = 623 wraps to new s[0]
    cnt = 0;
    for (int i = 454 + tid; i < 624; i += GL_THREADS) {
        uint32_t nxt = (i == 623) ? s[0] : s[i + 1];
        v[cnt++] = mt_mix(s[i], nxt, s[i - 227]);
    }
    __syncthreads();
    cnt = 0;
    for (int i = 454 + tid; i < 624; i += GL_THREADS) s[i] = v[cnt++];
    __syncthreads();
    for (int i = tid; i < 624; i += GL_THREADS) {
        uint32_t y = s[i];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        out[i] = y;
    }
    __syncthreads();
}

// Runs one partial permutation of n elements (first `need` final) consuming n-1 draws; perm lives in smem (u16).
__device__ void device_partial_perm(uint32_t *s, uint32_t *outbuf, int &avail_pos, uint16_t *perm, int n, int need) {
    for (int i = threadIdx.x; i < n; i += GL_THREADS) perm[i] = (uint16_t)i;
    __syncthreads();
    int draws = n > 0 ? n - 1 : 0;
    int i = 0;
    while (i < draws) {
        if (avail_pos >= 624) {  // uniform across the CTA
            mt_regen(s, outbuf);
            avail_pos = 0;
        }
        int take = min(624 - avail_pos, draws - i);
        if (threadIdx.x == 0) {
            for (int j = 0; j < take; j++) {
                int ii = i + j;
                if (ii >= need) break;
                uint32_t z = outbuf[avail_pos + j] % (uint32_t)(n - ii);
                uint16_t t = perm[ii];
                perm[ii] = perm[ii + z];
                perm[ii + z] = t;
            }
        }
        avail_pos += take;
        i += take;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(GL_THREADS) greenlist_build_kernel(
    int64_t V, int split, int seed_strategy, uint64_t salt, const int64_t *__restrict__ alive, int n_alive,
    const int64_t *__restrict__ dead, int n_dead, int green, int n_ga, int n_gd, uint32_t *__restrict__ table) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *s = reinterpret_cast<uint32_t *>(smem_raw);  // 624
    uint32_t *outbuf = s + 624;                            // 624
    uint16_t *perm = reinterpret_cast<uint16_t *>(outbuf + 624);
    const int64_t words = (V + 31) / 32;
    uint32_t *row = table + (int64_t)blockIdx.x * words;

    uint64_t seed = (seed_strategy == WMAR_SEED_FIXED) ? 0ull : context_seed(salt, (uint64_t)blockIdx.x);
    if (threadIdx.x == 0) {
        uint32_t v = (uint32_t)(seed & 0xffffffffu);
        s[0] = v;
        for (int i = 1; i < 624; i++) {
            v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
            s[i] = v;
        }
    }
    for (int64_t w = threadIdx.x; w < words; w += GL_THREADS) row[w] = 0u;
    __syncthreads();
    int avail_pos = 624;
    if (split == WMAR_SPLIT_RANDOM) {
        int need = green < (int)V ? green : (int)V;
        device_partial_perm(s, outbuf, avail_pos, perm, (int)V, need);
        for (int i = threadIdx.x; i < need; i += GL_THREADS) {
            uint32_t id = perm[i];
            atomicOr(&row[id >> 5], 1u << (id & 31));
        }
        return;
    }
    device_partial_perm(s, outbuf, avail_pos, perm, n_alive, n_ga);
    for (int i = threadIdx.x; i < n_ga; i += GL_THREADS) {
        int64_t id = alive[perm[i]];
        if (id >= 0 && id < V) atomicOr(&row[id >> 5], 1u << (id & 31));
    }
    __syncthreads();
    device_partial_perm(s, outbuf, avail_pos, perm, n_dead, n_gd);
    for (int i = threadIdx.x; i < n_gd; i += GL_THREADS) {
        int64_t id = dead[perm[i]];
        if (id >= 0 && id < V) atomicOr(&row[id >> 5], 1u << (id & 31));
    }
}

}  // namespace

extern "C" {

int wmar_version(void) { return 100; }
const char *wmar_last_error(void) { return wmar::g_last_error; }
uint64_t wmar_launch_count(void) { return wmar::g_launches.load(); }

int wmar_greenlist_build_host(int64_t V, double gamma, int split, int seed_strategy, uint64_t salt,
                              const int64_t *alive, int64_t n_alive, const int64_t *dead, int64_t n_dead,
                              int64_t n_rows, uint32_t *table, int n_threads) {
    WMAR_REQUIRE(V > 0 && V <= (1 << 20), "vocab_size out of range");
    WMAR_REQUIRE(gamma >= 0.0 && gamma <= 1.0, "gamma must be in [0,1]");
    WMAR_REQUIRE(split == WMAR_SPLIT_RANDOM || split == WMAR_SPLIT_RANDOM_STRATIFIED, "unsupported split strategy");
    WMAR_REQUIRE(table != nullptr && n_rows > 0, "bad table");
    WMAR_REQUIRE(seed_strategy != WMAR_SEED_FIXED || n_rows == 1, "FIXED seeding has exactly one row");
    WMAR_REQUIRE(split == WMAR_SPLIT_RANDOM || (n_alive >= 0 && n_dead >= 0 && (alive || !n_alive) && (dead || !n_dead)),
                 "bad alive/dead lists");
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    if ((int64_t)n_threads > n_rows) n_threads = (int)n_rows;
    const int64_t words = (V + 31) / 32;
    auto work = [&](int tid) {
        std::vector<int32_t> pa, pd;
        for (int64_t r = tid; r < n_rows; r += n_threads) {
            uint64_t seed = seed_strategy == WMAR_SEED_FIXED ? 0ull : context_seed(salt, (uint64_t)r);
            host_row(V, gamma, split, seed, alive, n_alive, dead, n_dead, table + r * words, pa, pd);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    return WMAR_OK;
}

int wmar_greenlist_build_device(int64_t V, double gamma, int split, int seed_strategy, uint64_t salt,
                                const int64_t *d_alive, int64_t n_alive, const int64_t *d_dead, int64_t n_dead,
                                int64_t n_rows, uint32_t *d_table, void *stream) {
    WMAR_REQUIRE(V > 0 && V <= 65536, "device build supports vocab_size <= 65536 (u16 permutation in smem)");
    WMAR_REQUIRE(gamma >= 0.0 && gamma <= 1.0, "gamma must be in [0,1]");
    WMAR_REQUIRE(split == WMAR_SPLIT_RANDOM || split == WMAR_SPLIT_RANDOM_STRATIFIED, "unsupported split strategy");
    WMAR_REQUIRE(d_table != nullptr && n_rows > 0 && n_rows < (1ll << 31), "bad table");
    WMAR_REQUIRE(seed_strategy != WMAR_SEED_FIXED || n_rows == 1, "FIXED seeding has exactly one row");
    WMAR_REQUIRE(n_alive <= 65536 && n_dead <= 65536, "alive/dead lists too long");
    const int green = (int)((double)V * gamma);
    int n_ga = (int)((double)n_alive * gamma);
    int n_gd = green - n_ga;
    if (n_ga > n_alive) n_ga = (int)n_alive;
    if (n_gd > n_dead) n_gd = (int)n_dead;
    if (n_gd < 0) n_gd = 0;
    int64_t max_n = split == WMAR_SPLIT_RANDOM ? V : (n_alive > n_dead ? n_alive : n_dead);
    size_t smem = sizeof(uint32_t) * 1248 + sizeof(uint16_t) * (size_t)(max_n + 8);
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(greenlist_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    greenlist_build_kernel<<<(unsigned)n_rows, GL_THREADS, smem, as_stream(stream)>>>(
        V, split, seed_strategy, salt, d_alive, (int)n_alive, d_dead, (int)n_dead, green, n_ga, n_gd, d_table);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

}  // extern "C"
