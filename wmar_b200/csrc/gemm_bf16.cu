// Launchers of the bf16-weight skinny GEMM (gemm_bf16.cuh).
#include "gemm_bf16.cuh"

namespace wmar {

template <int PRO, int EPI>
static int launch_bf16_t(const Bf16GemmArgs &a, cudaStream_t stream) {
    static int n_sms = 0;
    if (!n_sms) {
        int dev = 0;
        WMAR_CUDA_CHECK(cudaGetDevice(&dev));
        WMAR_CUDA_CHECK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int items = (a.N / GEMM_NT) * a.splits;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(items < 2 * n_sms ? items : 2 * n_sms), 1, 1);   // persistent: two CTAs per SM
    static bool configured = false;
    if (!configured) {
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(skinny_gemm_bf16_kernel<PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
        configured = true;
    }
    cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, skinny_gemm_bf16_kernel<PRO, EPI>, a));
    g_launches.fetch_add(1);
    return WMAR_OK;
}

int launch_skinny_gemm_bf16(int pro, int epi, const Bf16GemmArgs &a, cudaStream_t stream) {
    WMAR_REQUIRE(a.N % GEMM_NT == 0, "N must be a multiple of 64");
    WMAR_REQUIRE(a.splits >= 1 && a.K % (a.splits * BG_KI) == 0, "K must be a multiple of splits*32");
    WMAR_REQUIRE(a.splits == 1 || (a.ws != nullptr && a.counters != nullptr), "split-K needs a workspace");
    WMAR_REQUIRE(a.ldx % 4 == 0 && a.ldy % 4 == 0 && (a.K % 8) == 0, "row strides must be multiples of 4 floats");
#define WMAR_CASE(P, E) \
    if (pro == P && epi == E) return launch_bf16_t<P, E>(a, stream);
    WMAR_CASE(BPRO_NONE, BEPI_STORE)
    WMAR_CASE(BPRO_NONE, BEPI_RESID)
    WMAR_CASE(BPRO_RMS, BEPI_STORE)
    WMAR_CASE(BPRO_RMS, BEPI_STORE_F32)
    WMAR_CASE(BPRO_SWIGLU, BEPI_RESID)
    WMAR_CASE(BPRO_RMS, BEPI_SWIGLU)
    WMAR_CASE(BPRO_NONE, BEPI_STORE_F32)
#undef WMAR_CASE
    return set_error(WMAR_ERR_INVALID, "unsupported bf16 GEMM prologue/epilogue combination%s%s");
}

int pick_splits_bf16(int N, int K, int n_sms) {
    // the persistent grid has 2 * n_sms CTAs: pick the smallest split count that fills its waves to >= 85 % while every
    // item keeps >= 16 chunks of 32 k (two pipelined load rounds per warp)
    const int tiles = N / GEMM_NT, G = 2 * n_sms;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 64; s++) {
        if (K % (s * BG_KI) != 0) continue;
        if (K / (s * BG_KI) < 16 && s > 1) break;
        const int items = tiles * s;
        const int waves = (items + G - 1) / G;
        const double eff = (double)items / ((double)waves * G);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
        if (eff >= 0.85) break;
    }
    return best;
}

}  // namespace wmar

using namespace wmar;

namespace {
float *g_ws = nullptr;
unsigned *g_counters = nullptr;
size_t g_ws_bytes = 0, g_counter_n = 0;
}  // namespace

/* Stand-alone entry (unit tests): y[16][N] = bf16(x[16][K] . W[N][K]^T), W bf16, x / y fp32 (x is rounded to bf16). */
extern "C" int wmar_skinny_gemm_bf16(const float *d_x, const void *d_w_bf16, float *d_y, int64_t N, int64_t K, int split_k,
                                     void *stream) {
    WMAR_REQUIRE(d_x && d_w_bf16 && d_y && N > 0 && K > 0, "bad arguments");
    WMAR_REQUIRE(N % GEMM_NT == 0 && K % BG_KI == 0, "N % 64 == 0 and K % 32 == 0 required");
    int dev = 0, sms = 148;
    WMAR_CUDA_CHECK(cudaGetDevice(&dev));
    WMAR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int splits = split_k > 0 ? split_k : pick_splits_bf16((int)N, (int)K, sms);
    const size_t need = (size_t)(N / GEMM_NT) * splits * GEMM_M * GEMM_NT * sizeof(float);
    if (need > g_ws_bytes) {
        if (g_ws) cudaFree(g_ws);
        WMAR_CUDA_CHECK(cudaMalloc(&g_ws, need));
        g_ws_bytes = need;
    }
    if ((size_t)(N / GEMM_NT) > g_counter_n) {
        if (g_counters) cudaFree(g_counters);
        WMAR_CUDA_CHECK(cudaMalloc(&g_counters, sizeof(unsigned) * (size_t)(N / GEMM_NT)));
        WMAR_CUDA_CHECK(cudaMemset(g_counters, 0, sizeof(unsigned) * (size_t)(N / GEMM_NT)));
        g_counter_n = (size_t)(N / GEMM_NT);
    }
    Bf16GemmArgs a{};
    a.X = d_x; a.ldx = (int)K; a.W = reinterpret_cast<const __nv_bfloat16 *>(d_w_bf16); a.Y = d_y; a.ldy = (int)N;
    a.N = (int)N; a.K = (int)K; a.splits = splits; a.ws = g_ws; a.counters = g_counters;
    return launch_skinny_gemm_bf16(BPRO_NONE, BEPI_STORE, a, as_stream(stream));
}
