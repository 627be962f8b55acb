// Host side of the tcgen05 skinny GEMM: TMA tensor maps for the weight matrices (cached per pointer), split
// selection, launch through cudaLaunchKernelEx with programmatic dependent launch.
#include <mutex>
#include <unordered_map>

#include "gemm_tc.cuh"

namespace wmar {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct MapKey {
    const void *w; int N, K;
    bool operator==(const MapKey &o) const { return w == o.w && N == o.N && K == o.K; }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        return std::hash<const void *>()(k.w) ^ (std::hash<int>()(k.N) * 1315423911u) ^ (std::hash<int>()(k.K) * 2654435761u);
    }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// W[N][K] fp32 row-major -> 2-D map (K innermost), box 32 x 128, 128-byte swizzle
int weight_map(const float *W, int N, int K, CUtensorMap *out) {
    std::lock_guard<std::mutex> lk(g_map_mu);
    MapKey key{W, N, K};
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return WMAR_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver%s%s");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_KC, (cuuint32_t)TC_TILE_N};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(W), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled failed%s%s");
    g_maps.emplace(key, m);
    *out = m;
    return WMAR_OK;
}

int g_tc_dbg = 0;
int g_pdl = -1;
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char *e = getenv("WMAR_PDL");
        g_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pdl == 1;
}

template <int PRO, int EPI>
int launch_t(const CUtensorMap &map, const GemmArgs &a, int splits, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             TC_SMEM_ALLOC));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(a.N / TC_TILE_N), (unsigned)splits, 1);
    cfg.blockDim = dim3(TC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = TC_SMEM_ALLOC;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    TcExtra ex{splits, g_tc_dbg};
    WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<PRO, EPI>, map, a, ex));
    g_launches.fetch_add(1);
    return WMAR_OK;
}

}  // namespace

void tc_gemm_set_pdl(int on) { g_pdl = on ? 1 : 0; }
void tc_gemm_set_dbg(int bits) { g_tc_dbg = bits; }

int tc_weight_map(const float *W, int N, int K, CUtensorMap *out) { return weight_map(W, N, K, out); }

// W[N][K] fp32 row-major -> 2-D map (K innermost), box 16 x 64, no swizzle: one warp iteration of skinny_gemm_kernel
int tc_weight_map_skinny(const float *W, int N, int K, CUtensorMap *out) {
    std::lock_guard<std::mutex> lk(g_map_mu);
    MapKey key{W, -N, K};                       // negative N: skinny-GEMM entry of the shared cache
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return WMAR_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver%s%s");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    cuuint32_t box[2] = {16, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(W), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled (skinny GEMM weights) failed%s%s");
    g_maps.emplace(key, m);
    *out = m;
    return WMAR_OK;
}

// W[N][K] bf16 row-major -> 2-D map (K innermost), box 64 x 128 (128-byte rows), 128-byte swizzle (conv_tc.cuh, bf16x3)
int tc_weight_map_bf16(const void *W, int N, int K, CUtensorMap *out) {
    std::lock_guard<std::mutex> lk(g_map_mu);
    MapKey key{W, N, -K};                       // negative K: bf16 entry of the shared cache
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return WMAR_OK; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver%s%s");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(W), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled (bf16 weights) failed%s%s");
    g_maps.emplace(key, m);
    *out = m;
    return WMAR_OK;
}

// NHWC fp32 activations [B][H][W][C] -> 4-D map (C innermost), box [1][bh][bw][32], 128-byte swizzle, zero fill outside.
// estride = 2: every other pixel in x and y (the stride-2 convs): the box spans 2 bw x 2 bh source pixels, of which the
// TMA unit loads ceil(box / stride) = bw x bh.
int tc_nhwc_map(const float *base, int B, int H, int W, int C, int bw, int bh, CUtensorMap *out, int estride) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver%s%s");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(WMAR_ERR_CUDA, "cuTensorMapEncodeTiled (NHWC) failed%s%s");
    return WMAR_OK;
}
bool tc_available() { return encode_fn() != nullptr; }

void tc_gemm_forget_maps() {
    std::lock_guard<std::mutex> lk(g_map_mu);
    g_maps.clear();
}

bool tc_gemm_eligible(const GemmArgs &a) {
    return a.N % TC_TILE_N == 0 && a.K % TC_KC == 0 && a.K >= 2 * TC_KC && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0 &&
           ((size_t)a.K * sizeof(float)) % 16 == 0 && a.ldx % 4 == 0 && encode_fn() != nullptr;
}

int tc_pick_splits(int N, int K, int n_sms) {
    const int tiles = N / TC_TILE_N, C = K / TC_KC;
    int s = n_sms / (tiles > 0 ? tiles : 1);
    if (s > C / 2) s = C / 2;   // at least two 16 KB chunks per CTA
    if (s < 1) s = 1;
    return s;
}

int launch_tc_gemm(int pro, int epi, const GemmArgs &a, cudaStream_t stream) {
    int dev = 0, sms = 148;
    WMAR_CUDA_CHECK(cudaGetDevice(&dev));
    static int cached_sms = 0;
    if (!cached_sms) { WMAR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)); cached_sms = sms; }
    sms = cached_sms;
    const int splits = tc_pick_splits(a.N, a.K, sms);
    WMAR_REQUIRE(splits == 1 || (a.ws != nullptr && a.counters != nullptr), "split-K needs a workspace");
    CUtensorMap map;
    int rc = weight_map(a.W, a.N, a.K, &map);
    if (rc) return rc;
#define WMAR_CASE(P, E) \
    if (pro == P && epi == E) return launch_t<P, E>(map, a, splits, stream);
    WMAR_CASE(PRO_NONE, EPI_STORE)
    WMAR_CASE(PRO_NONE, EPI_RESID)
    WMAR_CASE(PRO_NONE, EPI_GATE_RESID)
    WMAR_CASE(PRO_NONE, EPI_GELU)
    WMAR_CASE(PRO_LN, EPI_STORE)
    WMAR_CASE(PRO_LN, EPI_GELU)
    WMAR_CASE(PRO_ADALN, EPI_STORE)
    WMAR_CASE(PRO_ADALN, EPI_GELU)
#undef WMAR_CASE
    return set_error(WMAR_ERR_INVALID, "unsupported GEMM prologue/epilogue combination%s%s");
}

}  // namespace wmar
