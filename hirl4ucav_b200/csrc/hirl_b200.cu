// hirl_b200.cu -- kernels' entry points and the C ABI of libhirl_b200.so (include/hirl_b200.h).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "env.cuh"
#include "update.cuh"
#include "cluster_update.cuh"
#include <condition_variable>
#include <mutex>
#include <thread>
#include "tcc_update.cuh"
#include "host_rng.h"

using namespace hirl;

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// The persistent fused update: k_steps learn() calls, one launch, team = whole grid.
__global__ void __launch_bounds__(THREADS, 1) hirl_update_kernel(const __grid_constant__ LaunchCtx L) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    // the context is read all over the (non-inlined) stage functions: keep it in shared memory,
    // a generic-address load from the kernel-parameter window costs a global-memory round trip
    __shared__ AgentCtx ctx;
    {
        const int *src = reinterpret_cast<const int *>(&L.agent);
        int *dst = reinterpret_cast<int *>(&ctx);
        for (int i = threadIdx.x; i < (int)(sizeof(AgentCtx) / sizeof(int)); i += THREADS) dst[i] = src[i];
    }
    __syncthreads();
    __shared__ tc::Barriers tc_bars;
    __shared__ uint32_t tmem_slot;
    const bool any_tc = ctx.use_tc || ctx.tc_g1;
    if (any_tc) {
        if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot);
        if (threadIdx.x == 32) tc::barriers_init(&tc_bars);
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    Team tm;
    tm.rank = blockIdx.x; tm.size = gridDim.x; tm.flags = L.bar_flags;
    tm.gen = ld_relaxed_u32(tm.flags + tm.rank * FLAG_STRIDE);   // flags keep counting across launches: resume from the generation this CTA reached
    tm.tmem_base = any_tc ? tmem_slot : 0u;
    tm.tc_smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    tm.tc_bars = &tc_bars;
    // fused data-parallel mode: the generations of this launch's exchange flags continue from the calls executed so far
    const unsigned dp_base = ctx.dp_world > 0 ? ld_relaxed_u32(ctx.dp_gen_ctr) : 0u;
    for (int step = 0; step < L.k_steps; ++step)
        if (learn_step(ctx, L.step0 + step, tm, smem, step == 0 ? L.stop_after : 0, L.prof, L.first_stage, L.last_stage, L.adam_phase, dp_base)) break;
    // (every CTA read the counter before the first grid barrier of the launch; CTA 0 leaves after the last one)
    if (ctx.dp_world > 0 && blockIdx.x == 0 && threadIdx.x == 0) *ctx.dp_gen_ctr = dp_base + (unsigned)L.k_steps;
    if (any_tc) {
        __syncthreads();
        if (threadIdx.x < 32) tc::tmem_dealloc(tm.tmem_base);
    }
}

// Ensemble form (BASELINE config 4): the grid is cut into teams of team_size CTAs; every team runs
// whole replicas (k_steps learn() calls each) with team-local barriers, so replicas never wait for
// each other and no CTA idles while another replica has tiles left.
__global__ void __launch_bounds__(THREADS, 1) hirl_ensemble_kernel(const __grid_constant__ EnsembleLaunch L) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    __shared__ AgentCtx ctx;
    const int team = blockIdx.x / L.team_size;
    if (team >= L.n_teams) return;
    __shared__ tc::Barriers tc_bars;
    __shared__ uint32_t tmem_slot;
    const bool use_tc = L.agents[0].use_tc != 0 || L.agents[0].tc_g1 != 0;   // any tensor-core tile in this launch
    if (use_tc) {
        if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot);
        if (threadIdx.x == 32) tc::barriers_init(&tc_bars);
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    Team tm;
    tm.rank = blockIdx.x % L.team_size; tm.size = L.team_size;
    tm.flags = L.bar_flags + (size_t)team * L.team_size * FLAG_STRIDE;
    tm.gen = ld_relaxed_u32(tm.flags + tm.rank * FLAG_STRIDE);
    tm.tmem_base = use_tc ? tmem_slot : 0u;
    tm.tc_smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
    tm.tc_bars = &tc_bars;
    for (int r = team; r < L.replicas; r += L.n_teams) {
        __syncthreads();
        const int *src = reinterpret_cast<const int *>(L.agents + r);
        int *dst = reinterpret_cast<int *>(&ctx);
        for (int i = threadIdx.x; i < (int)(sizeof(AgentCtx) / sizeof(int)); i += THREADS) dst[i] = src[i];
        __syncthreads();
        for (int step = 0; step < L.k_steps; ++step) learn_step(ctx, step, tm, smem, 0, nullptr);
    }
    if (use_tc) {
        __syncthreads();
        if (threadIdx.x < 32) tc::tmem_dealloc(tm.tmem_base);
    }
}

// buffer.py:38-48 + HIRL.py:229-233 as a standalone op (the fused kernel gathers in stage R0)
__global__ void gather_kernel(const float *__restrict__ online, const float *__restrict__ expert, const int32_t *__restrict__ idx_on,
                              int n_on, const int32_t *__restrict__ idx_ex, int n_ex, float *__restrict__ out) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n_on + n_ex; b += gridDim.x * wpb) {
        const float *src = b < n_on ? online + (size_t)idx_on[b] * ROW : expert + (size_t)idx_ex[b - n_on] * ROW;
        out[(size_t)b * ROW + lane] = __ldg(src + lane);
    }
}
__global__ void gather_bc_kernel(const float *__restrict__ tab, const int32_t *__restrict__ idx, int n, float *__restrict__ out_s,
                                 float *__restrict__ out_a) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < n; b += gridDim.x * wpb) {
        const float *src = tab + (size_t)idx[b] * BCW;
        if (lane < S) out_s[(size_t)b * S + lane] = __ldg(src + lane);
        else if (lane < SA) out_a[(size_t)b * A + lane - S] = __ldg(src + lane);
    }
}

// HIRL.py:192-212: actor forward for n rows, one CTA per row (latency path; no noise here)
__global__ void __launch_bounds__(THREADS) act_kernel(const float *net, const float *__restrict__ states, int n, float *__restrict__ actions,
                                                       float slope, const float *__restrict__ noise = nullptr, int noise_stride = 0) {
    __shared__ __align__(16) float h1[H1];
    __shared__ __align__(16) float z2[H2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int b = blockIdx.x; b < n; b += gridDim.x) {
        if (warp == 0) {
            const float w = lane < S ? states[(size_t)b * S + lane] : 0.f;
            float x[S];
#pragma unroll
            for (int k = 0; k < S; ++k) x[k] = __shfl_sync(0xffffffffu, w, k);
            l1_row<ActorL>(net, x, slope, lane, h1, nullptr, nullptr);
        }
        __syncthreads();
        float hreg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hreg[j] = h1[lane * 8 + j];
        for (int i = 0; i < H2 / WARPS; ++i) {
            const int col = warp * (H2 / WARPS) + i;
            const float *wr = net + ActorL::W2 + (size_t)col * H1 + lane * 8;
            const float4 w0 = ld4(wr), w1 = ld4(wr + 4);
            float p = hreg[0] * w0.x;
            p = fmaf(hreg[1], w0.y, p); p = fmaf(hreg[2], w0.z, p); p = fmaf(hreg[3], w0.w, p);
            p = fmaf(hreg[4], w1.x, p); p = fmaf(hreg[5], w1.y, p); p = fmaf(hreg[6], w1.z, p); p = fmaf(hreg[7], w1.w, p);
            p = warp_sum(p);
            if (lane == 0) z2[col] = p + net[ActorL::B2 + col];
        }
        __syncthreads();
        if (warp == 0) {
            float z[16], xh[16], h[16], rstd, out[A];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = ld4(z2 + 4 * lane + 128 * j);
                z[4 * j] = t.x; z[4 * j + 1] = t.y; z[4 * j + 2] = t.z; z[4 * j + 3] = t.w;
            }
            head_forward<ActorL>(net, z, slope, lane, xh, h, rstd, out);
            if (lane < A) {
                float a = tanhf(out[lane]);
                if (noise) a = fminf(fmaxf(a + noise[(size_t)b * noise_stride + lane], -1.f), 1.f);   // HIRL.py:196-198
                actions[(size_t)b * A + lane] = a;
            }
        }
        __syncthreads();
    }
}

// buffer.py:20-36 for n transitions at once: one warp packs one 128-byte replay row
__global__ void store_transitions_kernel(float *__restrict__ ring, int capacity, int first_slot, int n, const float *__restrict__ s,
                                         const float *__restrict__ a, const float *__restrict__ s2, const float *__restrict__ reward,
                                         const int32_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        float v;
        if (lane < S) v = s[(size_t)i * S + lane];
        else if (lane < SA) v = a[(size_t)i * A + lane - S];
        else if (lane < SA + S) v = s2[(size_t)i * S + lane - SA];
        else if (lane == 30) v = reward[i];
        else v = (flags[i] & 1) ? 1.f : 0.f;
        ring[(size_t)((first_slot + i) % capacity) * ROW + lane] = v;
    }
}

// register-resident FFMA chains: the FP32 roofline denominator (see hirl_measure_fp32_peak)
__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(HIRL_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// ---- NCCL, resolved at run time (data-parallel mode, see launch_update_dp) ------------------------------------------------
#include <dlfcn.h>
namespace {
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_uid_fn)(void *);
typedef int (*nccl_init_fn)(void **, int, const void *, int);   // ncclCommInitRank(comm*, nranks, ncclUniqueId BY VALUE, rank): see nccl_init_rank below
typedef int (*nccl_destroy_fn)(void *);
typedef const char *(*nccl_errstr_fn)(int);
typedef int (*nccl_group_fn)(void);
struct NcclUid { char internal[128]; };
typedef int (*nccl_init_byval_fn)(void **, int, NcclUid, int);
struct NcclApi {
    void *lib = nullptr;
    nccl_allreduce_fn allreduce = nullptr;
    nccl_uid_fn unique_id = nullptr;
    nccl_init_byval_fn init_rank = nullptr;
    nccl_destroy_fn destroy = nullptr;
    nccl_errstr_fn errstr = nullptr;
    nccl_group_fn group_start = nullptr, group_end = nullptr;
} g_nccl;
int nccl_load() {
    if (g_nccl.lib) return 0;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(HIRL_E_UNSUPPORTED, std::string("libnccl.so.2 not found: ") + dlerror());
    g_nccl.allreduce = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
    g_nccl.unique_id = (nccl_uid_fn)dlsym(lib, "ncclGetUniqueId");
    g_nccl.init_rank = (nccl_init_byval_fn)dlsym(lib, "ncclCommInitRank");
    g_nccl.destroy = (nccl_destroy_fn)dlsym(lib, "ncclCommDestroy");
    g_nccl.errstr = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
    g_nccl.group_start = (nccl_group_fn)dlsym(lib, "ncclGroupStart");
    g_nccl.group_end = (nccl_group_fn)dlsym(lib, "ncclGroupEnd");
    if (!g_nccl.allreduce || !g_nccl.unique_id || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.errstr || !g_nccl.group_start || !g_nccl.group_end)
        return fail(HIRL_E_UNSUPPORTED, "libnccl.so.2 lacks an expected symbol");
    g_nccl.lib = lib;
    return 0;
}
constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)
}  // namespace
#define NC(call)                                                                                                    \
    do {                                                                                                            \
        int r_ = (call);                                                                                            \
        if (r_ != 0) return fail(HIRL_E_CUDA, std::string(#call) + ": " + (g_nccl.errstr ? g_nccl.errstr(r_) : "nccl error")); \
    } while (0)

struct hirl_agent {
    hirl_config_t cfg;
    int sm_count = 0, grid = 0;
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;
    std::vector<AgentCtx> ctxs;          // one per replica (ctxs[r].sc = scratch of team r % n_teams)
    std::vector<float *> online, expert;  // per-replica replay tables
    AgentCtx *d_ctxs = nullptr;           // device copy of ctxs (ensemble launches)
    int n_teams = 1, team_size = 0;       // replicas > 1: teams of team_size CTAs
    unsigned *d_bar = nullptr;
    float *bctab = nullptr;               // expert (s,a) table, shared read-only by all replicas
    hirl_step_t *d_steps = nullptr;
    int32_t *d_ion = nullptr, *d_iex = nullptr, *d_ibc = nullptr;
    float *d_out = nullptr;
    int stop_after = 0, last_launches = 0;
    long long *d_prof = nullptr;
    bool prof_on = false;
    bool zero_copy = false;      // HIRL_ZERO_COPY=1: kernel reads the pinned host streams over PCIe (measured slower, see below)
    // one-call CUDA graph of the per-call sequence (4 H2D copies, cooperative launch, D2H): a single cudaGraphLaunch
    // instead of six API calls; rebuilt when the caller's staging pointers change, disabled if capture is refused
    float *d_act_s = nullptr, *d_act_a = nullptr;   // hirl_act staging (states in, actions out), act_cap rows
    int act_cap = 0;
    bool use_graph = true;
    cudaGraphExec_t graph_exec = nullptr;
    const void *graph_key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // per-call path: the step record and the three index rows travel as ONE packed H2D copy (library-owned pinned block)
    char *h_pack = nullptr, *d_pack = nullptr;
    size_t pack_bytes = 0;
    size_t scratch_elems[32];
    bool coop_ok = true;         // cooperative attribute accepted for cluster launches
    int ncl = 0;                 // cluster kernel: co-resident clusters of 8 CTAs used per launch
    int tcc_ncl = 0;             // tensor-core cluster kernel: clusters per launch
    // data-parallel large-batch mode (hirl_allreduce_init / hirl_allreduce_attach)
    void *nccl_comm = nullptr;
    bool dp_own_comm = false;
    int dp_world = 0, dp_rank = 0;
    float *d_dp_count = nullptr;
    std::vector<int> dp_do_actor, dp_mode;   // host copy of the step records' control words of the launch being enqueued
    // fused variant: in-kernel all-reduce over IPC-mapped peer memory
    bool dp_fused = false;
    float *d_dp_xchg = nullptr;
    std::vector<void *> dp_opened;           // peer mappings to close
};

template <class T>
static int dev_alloc(hirl_agent *h, T **p, size_t n, bool zero = true) {
    void *q = nullptr;
    CU(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    if (zero) CU(cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T)));
    h->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return 0;
}

// ---- state_dict <-> internal layout --------------------------------------------------------
// state_dict order of one tower: W1 (H1,K0) row-major, b1, g1, be1, W2 (H2,H1), b2, g2, be2, W3 (NOUT,H2), b3
template <class L>
static void tower_to_internal(const float *src, float *dst) {
    std::fill(dst, dst + L::SIZE, 0.f);
    const float *p = src;
    for (int n = 0; n < H1; ++n)
        for (int k = 0; k < L::k0; ++k) dst[L::W1T + k * H1 + n] = p[n * L::k0 + k];
    p += H1 * L::k0;
    std::memcpy(dst + L::B1, p, H1 * 4); p += H1;
    std::memcpy(dst + L::G1, p, H1 * 4); p += H1;
    std::memcpy(dst + L::BE1, p, H1 * 4); p += H1;
    std::memcpy(dst + L::W2, p, (size_t)H2 * H1 * 4); p += H2 * H1;
    std::memcpy(dst + L::B2, p, H2 * 4); p += H2;
    std::memcpy(dst + L::G2, p, H2 * 4); p += H2;
    std::memcpy(dst + L::BE2, p, H2 * 4); p += H2;
    std::memcpy(dst + L::W3, p, (size_t)L::nout * H2 * 4); p += L::nout * H2;
    std::memcpy(dst + L::B3, p, L::nout * 4);
}
template <class L>
static void tower_from_internal(const float *src, float *dst) {
    float *p = dst;
    for (int n = 0; n < H1; ++n)
        for (int k = 0; k < L::k0; ++k) p[n * L::k0 + k] = src[L::W1T + k * H1 + n];
    p += H1 * L::k0;
    std::memcpy(p, src + L::B1, H1 * 4); p += H1;
    std::memcpy(p, src + L::G1, H1 * 4); p += H1;
    std::memcpy(p, src + L::BE1, H1 * 4); p += H1;
    std::memcpy(p, src + L::W2, (size_t)H2 * H1 * 4); p += H2 * H1;
    std::memcpy(p, src + L::B2, H2 * 4); p += H2;
    std::memcpy(p, src + L::G2, H2 * 4); p += H2;
    std::memcpy(p, src + L::BE2, H2 * 4); p += H2;
    std::memcpy(p, src + L::W3, (size_t)L::nout * H2 * 4); p += L::nout * H2;
    std::memcpy(p, src + L::B3, L::nout * 4);
}
template <class L>
constexpr size_t tower_flat() { return (size_t)H1 * L::k0 + 3 * H1 + (size_t)H2 * H1 + 3 * H2 + (size_t)L::nout * H2 + L::nout; }
static_assert(tower_flat<ActorL>() == HIRL_ACTOR_PARAMS, "actor flat size");
static_assert(2 * tower_flat<CriticL>() == HIRL_CRITIC_PARAMS, "critic flat size");

static bool is_critic(int net) { return net == HIRL_NET_CRITIC || net == HIRL_NET_TARGET_CRITIC; }
static int net_ptr(hirl_agent *h, int replica, int net, int kind /*0 w,1 g,2 m,3 v*/, float **out) {
    if (replica < 0 || replica >= (int)h->ctxs.size()) return fail(HIRL_E_INVALID, "replica index out of range");
    AgentCtx &c = h->ctxs[replica];
    float *p = nullptr;
    if (kind == 0) p = net == HIRL_NET_ACTOR ? c.actor : net == HIRL_NET_TARGET_ACTOR ? c.tactor : net == HIRL_NET_CRITIC ? c.critic
                   : net == HIRL_NET_TARGET_CRITIC ? c.tcritic : net == HIRL_NET_BC_ACTOR ? c.bc : nullptr;
    else if (net == HIRL_NET_ACTOR) p = kind == 1 ? c.g_actor : kind == 2 ? c.m_actor : c.v_actor;
    else if (net == HIRL_NET_CRITIC) p = kind == 1 ? c.g_critic : kind == 2 ? c.m_critic : c.v_critic;
    if (!p) return fail(HIRL_E_INVALID, "bad network selector");
    *out = p;
    return 0;
}
static int flat_import(hirl_agent *h, int replica, int net, int kind, const float *src, size_t n) {
    float *d;
    if (int rc = net_ptr(h, replica, net, kind, &d)) return rc;
    const bool cr = is_critic(net);
    if (n != (cr ? (size_t)HIRL_CRITIC_PARAMS : (size_t)HIRL_ACTOR_PARAMS)) return fail(HIRL_E_INVALID, "flat parameter vector has the wrong length");
    std::vector<float> tmp(cr ? CRITIC_SIZE : ACTOR_SIZE);
    if (cr) {
        tower_to_internal<CriticL>(src, tmp.data());
        tower_to_internal<CriticL>(src + tower_flat<CriticL>(), tmp.data() + CriticL::SIZE);
    } else tower_to_internal<ActorL>(src, tmp.data());
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    // on the agent's stream: a blocking cudaMemcpy from pageable memory may return before the last staged chunk has landed,
    // and kernels launched on h->stream right afterwards (the image builder below) are not ordered behind the legacy stream
    CU(cudaMemcpyAsync(d, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (kind == 0 && h->ctxs[replica].use_tcc) {   // tensor-core operand images of W2 (csrc/tcc_update.cuh), rebuilt on the device
        AgentCtx &c = h->ctxs[replica];
        const int towers = cr ? 2 : 1;
        const int w2off = cr ? CriticL::W2 : ActorL::W2, tsize = cr ? CriticL::SIZE : ActorL::SIZE;
        float *img = net == HIRL_NET_ACTOR ? c.tcc.img_actor : net == HIRL_NET_TARGET_ACTOR ? c.tcc.img_tactor : net == HIRL_NET_BC_ACTOR ? c.tcc.img_bc
                     : net == HIRL_NET_CRITIC ? c.tcc.img_critic : c.tcc.img_tcritic;
        float *imgt = net == HIRL_NET_ACTOR ? c.tcc.imgt_actor : net == HIRL_NET_CRITIC ? c.tcc.imgt_critic : nullptr;
        for (int tw = 0; tw < towers; ++tw)
            tcc::tcc_images_kernel<<<128, 256, 0, h->stream>>>(d + (size_t)tw * tsize + w2off, img + (size_t)tw * tcc::IMG_FWD,
                                                                imgt ? imgt + (size_t)tw * tcc::IMG_T : nullptr);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    }
    if (kind == 0 && h->ctxs[replica].use_cluster) {   // padded W2 image ([H2][H1 + 4] per tower) streamed by the cluster kernel
        const int towers = cr ? 2 : 1;
        const int w2off = cr ? CriticL::W2 : ActorL::W2, tsize = cr ? CriticL::SIZE : ActorL::SIZE;
        std::vector<float> img((size_t)towers * H2 * cl::WLD, 0.f);
        for (int tw = 0; tw < towers; ++tw)
            for (int n = 0; n < H2; ++n)
                std::memcpy(&img[((size_t)tw * H2 + n) * cl::WLD], &tmp[(size_t)tw * tsize + w2off + (size_t)n * H1], H1 * 4);
        AgentCtx &c = h->ctxs[replica];
        float *dst = net == HIRL_NET_ACTOR ? c.w2p_actor : net == HIRL_NET_TARGET_ACTOR ? c.w2p_tactor : net == HIRL_NET_BC_ACTOR ? c.w2p_bc
                     : net == HIRL_NET_CRITIC ? c.w2p_critic : c.w2p_tcritic;
        CU(cudaMemcpy(dst, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
    }
    if (kind == 0 && (net == HIRL_NET_ACTOR || net == HIRL_NET_CRITIC)) {   // W2^T copies of the trainable nets
        const int towers = cr ? 2 : 1;
        const int w2off = cr ? CriticL::W2 : ActorL::W2, tsize = cr ? CriticL::SIZE : ActorL::SIZE;
        std::vector<float> t((size_t)towers * H1 * H2);
        for (int tw = 0; tw < towers; ++tw)
            for (int n = 0; n < H2; ++n)
                for (int k = 0; k < H1; ++k) t[(size_t)tw * H1 * H2 + (size_t)k * H2 + n] = tmp[(size_t)tw * tsize + w2off + (size_t)n * H1 + k];
        CU(cudaMemcpy(cr ? h->ctxs[replica].w2t_critic : h->ctxs[replica].w2t_actor, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}
static int flat_export(hirl_agent *h, int replica, int net, int kind, float *dst, size_t n) {
    float *d;
    if (int rc = net_ptr(h, replica, net, kind, &d)) return rc;
    const bool cr = is_critic(net);
    if (n != (cr ? (size_t)HIRL_CRITIC_PARAMS : (size_t)HIRL_ACTOR_PARAMS)) return fail(HIRL_E_INVALID, "flat parameter vector has the wrong length");
    std::vector<float> tmp(cr ? CRITIC_SIZE : ACTOR_SIZE);
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(tmp.data(), d, tmp.size() * 4, cudaMemcpyDeviceToHost));
    if (cr) {
        tower_from_internal<CriticL>(tmp.data(), dst);
        tower_from_internal<CriticL>(tmp.data() + CriticL::SIZE, dst + tower_flat<CriticL>());
    } else tower_from_internal<ActorL>(tmp.data(), dst);
    return 0;
}

// split-K choice.  A stage is a list of (tiles, cost) jobs handed out round-robin exactly as
// run_gemm_jobs does; pick the split of the splittable jobs that minimises the busiest CTA.
struct JobShape { int tiles; double cost; bool splittable; };
static double stage_makespan(const std::vector<JobShape> &jobs, int ns, int ctas) {
    std::vector<double> load(ctas, 0.0);
    int t = 0;
    for (const JobShape &j : jobs) {
        const int n = j.splittable ? j.tiles * ns : j.tiles;
        const double c = j.splittable ? j.cost / ns + 2.0 : j.cost + 2.0;  // +2: per-tile prologue/epilogue
        for (int i = 0; i < n; ++i, ++t) load[t % ctas] += c;
    }
    return *std::max_element(load.begin(), load.end());
}
static int pick_split(const std::vector<JobShape> &jobs, int ctas, int forced) {
    if (forced > 0) return std::min(forced, MAX_SPLIT);
    int best = 1;
    double best_t = 1e300;
    for (int ns = 1; ns <= MAX_SPLIT; ns *= 2) {
        const double t = stage_makespan(jobs, ns, ctas) * (1.0 + 0.03 * (ns - 1));  // partial-sum traffic penalty
        if (t < best_t - 1e-9) { best_t = t; best = ns; }
    }
    return best;
}

extern "C" {

int hirl_abi_version(void) { return HIRL_ABI_VERSION; }
const char *hirl_last_error(void) { return g_err.c_str(); }

int hirl_agent_create(const hirl_config_t *cfg, hirl_agent_t **out) {
    if (!cfg || !out) return fail(HIRL_E_INVALID, "null argument");
    if (cfg->batch < 1 || cfg->online_capacity < 1 || cfg->expert_capacity < 1 || cfg->bc_rows < 0 || cfg->max_steps < 1)
        return fail(HIRL_E_INVALID, "batch/capacities/max_steps must be positive");
    if (cfg->replicas < 1 || cfg->replicas > 4096) return fail(HIRL_E_INVALID, "replicas must be in [1, 4096]");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(HIRL_E_INVALID, "no such CUDA device");
    CU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(HIRL_E_UNSUPPORTED, "libhirl_b200 is built for sm_100a (B200) only");
    hirl_agent *h = new hirl_agent();
    h->cfg = *cfg;
    h->sm_count = prop.multiProcessorCount;
    CU(cudaFuncSetAttribute(hirl_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hirl_update_kernel, THREADS, SMEM_BYTES));
    if (occ < 1) return fail(HIRL_E_CUDA, "update kernel does not fit on an SM");
    h->grid = std::min(h->sm_count, MAX_CTAS) & ~1;  // even: the row stages split CTAs in two groups
    if (const char *g = getenv("HIRL_GRID")) h->grid = std::max(2, std::min(h->grid, atoi(g)) & ~1);
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (const char *z = getenv("HIRL_ZERO_COPY")) h->zero_copy = atoi(z) != 0;
    if (const char *z = getenv("HIRL_GRAPH")) h->use_graph = atoi(z) != 0;

    const int R = cfg->replicas;
    const size_t B = cfg->batch;
    int team_ctas = h->grid;  // CTAs that cooperate on one agent
    if (R > 1) {
        h->team_size = getenv("HIRL_TEAM") ? std::max(2, atoi(getenv("HIRL_TEAM")) & ~1) : 8;
        h->team_size = std::min(h->team_size, h->grid);
        h->n_teams = std::min(h->grid / h->team_size, R);
        team_ctas = h->team_size;
    }
    AgentCtx proto;
    std::memset(&proto, 0, sizeof(proto));
    proto.B = cfg->batch;
    proto.Bg = cfg->batch;
    proto.dp_count = nullptr;
    proto.dp_world = 0; proto.dp_rank = 0; proto.dp_gen_ctr = nullptr;
    for (int r = 0; r < 8; ++r) proto.dp_peer[r] = nullptr;
    proto.activation = cfg->activation;
    proto.gamma = cfg->gamma; proto.tau = cfg->tau; proto.noise_clamp = cfg->noise_clamp; proto.loss_lambda = cfg->loss_lambda;
    // split-K factors (unit = one 64x64 output tile over the layer's full K)
    const int forced = getenv("HIRL_NSPLIT") ? atoi(getenv("HIRL_NSPLIT")) : 0;
    const int tm = (cfg->batch + 63) / 64;
    const double kb = std::ceil(cfg->batch / 16.0);                  // k-tiles of a batch-long reduction
    const JobShape fwd{tm * 8, 16.0, true}, dx{tm * 4, 32.0, true};  // K=256 / K=512 in 16-wide k-tiles
    proto.ns[NS_G1] = pick_split({fwd, fwd, fwd}, team_ctas, forced);
    proto.ns[NS_G1_ACTOR] = pick_split({fwd, fwd, fwd, fwd, fwd, fwd}, team_ctas, forced);
    proto.ns[NS_G2] = pick_split({fwd, fwd}, team_ctas, forced);
    proto.ns[NS_G3DX] = pick_split({{32, kb, false}, {32, kb, false}, dx, dx}, team_ctas, forced);
    proto.ns[NS_G5] = pick_split({fwd, fwd}, team_ctas, forced);
    proto.ns[NS_G6] = pick_split({dx}, team_ctas, forced);
    proto.ns[NS_G7DX] = pick_split({{32, kb, false}, {32, kb, false}, dx, dx}, team_ctas, forced);
    // tensor-core path (tcgen05, 3xTF32, 128x64 tiles).  Measured: at batch 256 a single agent has at most one
    // tile per SM per stage and the FP32 64x64 tiles (twice as many, split-K) finish sooner; once a CTA has several
    // tiles per stage (large batch, or a team of 8 CTAs per replica in ensembles) the tensor-core tiles win 2.5-3.5x.
    proto.use_tc = (cfg->batch >= 512 || (R > 1 && cfg->batch >= 128)) ? 1 : 0;
    // Mid-size batches (single agent, 128 <= batch < 512): only stage G1 goes to the tensor cores.  Its 3-6 forward
    // passes are 16 tiles of 128x64x256 each -> one wave of 48-96 concurrent tcgen05 tiles (10.4 us measured) against
    // 24 us for the FP32 tiles; the other GEMM stages have too few tiles (G2, G5: 32) or would need the transposed
    // operand copies (dX, dW), where the FP32 tiles stay faster at this size.
    proto.tc_g1 = (!proto.use_tc && R == 1 && cfg->batch >= 128) ? 1 : 0;
    if (const char *e = getenv("HIRL_TC")) {   // 0: FP32 tiles only, 1: every 256<->512 contraction on tcgen05, 2: stage G1 only
        proto.use_tc = atoi(e) == 1;
        proto.tc_g1 = atoi(e) == 2;
    }
    // Cluster-resident kernel (csrc/cluster_update.cuh): single agent, FP32 path, selected with HIRL_CLUSTER=1.  Measured
    // on B200 (profiles/r2_cluster_timeline.txt) it needs 4 instead of 16 grid barriers per actor call and keeps the
    // activations on chip, but every one of its sequential sub-steps is latency-bound with 8 warps per SM: 132 us per
    // batch-256 step against 112 us for the stage-structured kernel, which therefore stays the default.
    proto.use_cluster = 0;
#ifndef HIRL_NO_CLUSTER   // development builds (-DHIRL_NO_CLUSTER) leave the two cluster-kernel instances out: half the compile time
    if (const char *e = getenv("HIRL_CLUSTER")) proto.use_cluster = (atoi(e) != 0 && R == 1) ? 1 : 0;
#endif
    proto.cluster_ri = 8;
    proto.Bpad = ((cfg->batch + 15) / 16) * 16;
    // Tensor-core cluster kernel (csrc/tcc_update.cuh): the default for a single HIRL agent (relu towers) at batch <= 256.
    // HIRL_TCC=0/1 overrides.  TD3 / BC step modes are not implemented there; those agents (leaky_relu towers, TD3.py / BC.py)
    // and larger batches / ensembles run on the stage kernel.
    proto.use_tcc = (R == 1 && cfg->batch <= 4 * tcc::RB && cfg->activation == HIRL_ACT_RELU && !proto.use_cluster) ? 1 : 0;
    if (const char *e = getenv("HIRL_TCC")) proto.use_tcc = (atoi(e) != 0 && R == 1 && cfg->batch <= 4 * tcc::RB && !proto.use_cluster) ? 1 : 0;
    if (proto.use_tcc && (getenv("HIRL_TC") || getenv("HIRL_NSPLIT") || getenv("HIRL_DW_SPLIT")) && !getenv("HIRL_TCC")) proto.use_tcc = 0;   // explicit stage-kernel knobs select the stage kernel
    if (proto.use_tcc) {
        CU(cudaFuncSetAttribute(tcc::hirl_tcc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcc::SMEM_BYTES));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(tcc::MAX_NCL * tcc::CL); lc.blockDim = dim3(tcc::TTHREADS); lc.dynamicSmemBytes = tcc::SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = tcc::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        int ncl = 0;
        CU(cudaOccupancyMaxActiveClusters(&ncl, (const void *)tcc::hirl_tcc_kernel, &lc));
        ncl = std::min(ncl, tcc::MAX_NCL);
        if (const char *e = getenv("HIRL_NCL")) ncl = std::min(ncl, std::max(1, atoi(e)));
        const int nrb = (cfg->batch + tcc::RB - 1) / tcc::RB;
        if (ncl < 3 * nrb) proto.use_tcc = 0;                        // every unit of a phase needs its own co-resident cluster
        h->tcc_ncl = ncl;
        if (proto.use_tcc) proto.use_tc = proto.tc_g1 = 0;
    }
#ifndef HIRL_NO_CLUSTER
    if (proto.use_cluster) {
        proto.use_tc = proto.tc_g1 = 0;
        // how many clusters of 8 CTAs are co-resident on this part (B200: 15 or 16, depends on the floor plan of the die)
        int ncl[2] = {0, 0};
        for (int v = 0; v < 2; ++v) {
            const void *fn = v ? (const void *)cl::hirl_cluster_kernel<9> : (const void *)cl::hirl_cluster_kernel<8>;
            const size_t sm = v ? cl::Ker<9>::SMEM_BYTES : cl::Ker<8>::SMEM_BYTES;
            CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(cl::MAX_CL * cl::CL); lc.blockDim = dim3(cl::CTHREADS); lc.dynamicSmemBytes = sm;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            CU(cudaOccupancyMaxActiveClusters(&ncl[v], fn, &lc));
            ncl[v] = std::min(ncl[v], cl::MAX_CL);
        }
        if (const char *e = getenv("HIRL_NCL")) { ncl[0] = std::min(ncl[0], std::max(1, atoi(e))); ncl[1] = std::min(ncl[1], std::max(1, atoi(e))); }
        if (ncl[0] < 1) return fail(HIRL_E_CUDA, "no co-resident cluster of 8 CTAs for the cluster kernel");
        // rows per block: 16, or 18 when that saves a whole round of row blocks (e.g. batch 256 on 15 clusters)
        const int r16 = ((cfg->batch + 15) / 16 + ncl[0] - 1) / ncl[0] * 16;
        const int r18 = ncl[1] >= 1 ? ((cfg->batch + 17) / 18 + ncl[1] - 1) / ncl[1] * 18 : 1 << 30;
        proto.cluster_ri = r18 < r16 ? 9 : 8;
        if (const char *e = getenv("HIRL_RI")) proto.cluster_ri = atoi(e) == 9 ? 9 : 8;
        h->ncl = proto.cluster_ri == 9 ? ncl[1] : ncl[0];
        const int rb = 2 * proto.cluster_ri;
        proto.Bpad = ((cfg->batch + rb - 1) / rb) * rb;
    }
#endif
    proto.BT = (cfg->batch + 31) & ~31;
    // dW2 reduces over the batch: split it so that a long reduction still spreads over the SMs (tensor-core tiles:
    // one split per 512 batch rows = one accumulation chunk; FP32 tiles: one per 1024 rows)
    proto.dw_split = std::max(1, std::min(8, proto.use_tc ? proto.BT / 512 : proto.BT / 1024));
    if (const char *e = getenv("HIRL_DW_SPLIT")) proto.dw_split = std::max(1, std::min(8, atoi(e)));
    if (proto.use_tc)
        for (int i = 0; i < NS_COUNT; ++i) proto.ns[i] = 1;
    if (proto.tc_g1) proto.ns[NS_G1] = proto.ns[NS_G1_ACTOR] = 1;   // tensor-core tiles take the full K (splitting the
                                                                    // 48 tiles of a critic-only call in two K halves
                                                                    // made G1 3 us shorter and the step no faster)
    proto.zsplit = 1;
    for (int i = 0; i < NS_COUNT; ++i) proto.zsplit = std::max(proto.zsplit, proto.ns[i]);

    int rc = 0;
    if ((rc = dev_alloc(h, &h->bctab, (size_t)std::max(cfg->bc_rows, 1) * BCW))) return rc;
    // scratch: one set per team
    std::vector<Scratch> team_sc(h->n_teams);
    size_t *se = h->scratch_elems;
    for (int t = 0; t < h->n_teams; ++t) {
        Scratch &s = team_sc[t];
#define SCR(field, id, n)                                        \
    if ((rc = dev_alloc(h, &s.field, (n)))) return rc;            \
    se[id] = (n);
        SCR(X, DB_X, B * ROW) SCR(XB, DB_XB, B * BCW)
        SCR(H1, DB_H1, (size_t)N_PASS * B * H1) SCR(XH1, DB_XH1, (size_t)N_BWD * B * H1) SCR(RS1, DB_RS1, (size_t)N_BWD * B)
        SCR(Z2, DB_Z2, (size_t)N_PASS * proto.zsplit * B * H2) SCR(DH1, DB_DH1, (size_t)N_BWD * proto.zsplit * B * H1)
        SCR(APRIME, DB_APRIME, B * A) SCR(PI, DB_PI, B * A) SCR(PIB, DB_PIB, B * A) SCR(ABC, DB_ABC, B * A) SCR(GPI, DB_GPI, B * A)
        SCR(Y, DB_Y, B) SCR(Q, DB_Q, 2 * B) SCR(TQ, DB_TQ, 2 * B) SCR(RLQ, DB_RLQ, B) SCR(SOFTQ, DB_SOFTQ, B)
        SCR(LP, DB_LP, (size_t)MAX_CTAS * LP_WORDS)
#undef SCR
        const int slots = R > 1 ? h->team_size : MAX_CTAS;
        if ((rc = dev_alloc(h, &s.GW2P, (size_t)2 * proto.dw_split * H2 * H1))) return rc;
        if ((rc = dev_alloc(h, &s.H1T, (size_t)4 * H1 * proto.BT))) return rc;
        if ((rc = dev_alloc(h, &s.DZ2T, (size_t)4 * H2 * proto.BT))) return rc;
        if ((rc = dev_alloc(h, &s.H1P, (size_t)N_PASS * proto.Bpad * cl::WLD))) return rc;
        if ((rc = dev_alloc(h, &s.DZ2S, (size_t)4 * cl::CL * proto.Bpad * cl::NS))) return rc;
        if ((rc = dev_alloc(h, &s.DZ1, (size_t)4 * B * H1))) return rc;
        if ((rc = dev_alloc(h, &s.DY1, (size_t)4 * B * H1))) return rc;
        if ((rc = dev_alloc(h, &s.XH2, (size_t)2 * proto.Bpad * H2))) return rc;
        if ((rc = dev_alloc(h, &s.RS2, (size_t)2 * proto.Bpad))) return rc;
        if ((rc = dev_alloc(h, &s.PART2C, (size_t)cl::MAX_CL * cl::PC_WORDS))) return rc;
        if ((rc = dev_alloc(h, &s.PART2A, (size_t)cl::MAX_CL * cl::PA_WORDS))) return rc;
        if ((rc = dev_alloc(h, &s.PART_C, (size_t)slots * 2 * CriticL::SMALL))) return rc;
        if ((rc = dev_alloc(h, &s.PART_A, (size_t)slots * ActorL::SMALL))) return rc;
    }
    // per-replica state: parameter arena, replay tables, carried scalars
    h->ctxs.assign(R, proto);
    h->online.assign(R, nullptr);
    h->expert.assign(R, nullptr);
    const size_t arena_n = 6 * (size_t)ACTOR_SIZE + 5 * (size_t)CRITIC_SIZE;
    for (int r = 0; r < R; ++r) {
        AgentCtx &c = h->ctxs[r];
        float *arena = nullptr;
        if ((rc = dev_alloc(h, &arena, arena_n))) return rc;
        float *p = arena;
        c.actor = p; p += ACTOR_SIZE; c.tactor = p; p += ACTOR_SIZE; c.bc = p; p += ACTOR_SIZE;
        c.g_actor = p; p += ACTOR_SIZE; c.m_actor = p; p += ACTOR_SIZE; c.v_actor = p; p += ACTOR_SIZE;
        c.critic = p; p += CRITIC_SIZE; c.tcritic = p; p += CRITIC_SIZE;
        c.g_critic = p; p += CRITIC_SIZE; c.m_critic = p; p += CRITIC_SIZE; c.v_critic = p; p += CRITIC_SIZE;
        if ((rc = dev_alloc(h, &h->online[r], (size_t)cfg->online_capacity * ROW))) return rc;
        if ((rc = dev_alloc(h, &h->expert[r], (size_t)cfg->expert_capacity * ROW))) return rc;
        c.online = h->online[r]; c.expert = h->expert[r]; c.bctab = h->bctab;
        c.sc = team_sc[r % h->n_teams];
        if ((rc = dev_alloc(h, &c.carry, 8))) return rc;
        if ((rc = dev_alloc(h, &c.w2t_actor, (size_t)H1 * H2))) return rc;
        if ((rc = dev_alloc(h, &c.w2t_critic, (size_t)2 * H1 * H2))) return rc;
        if (proto.use_tcc) {
            TccPtrs &t = c.tcc;
            if ((rc = dev_alloc(h, &t.img_actor, tcc::IMG_FWD))) return rc;
            if ((rc = dev_alloc(h, &t.img_tactor, tcc::IMG_FWD))) return rc;
            if ((rc = dev_alloc(h, &t.img_bc, tcc::IMG_FWD))) return rc;
            if ((rc = dev_alloc(h, &t.img_critic, 2 * tcc::IMG_FWD))) return rc;
            if ((rc = dev_alloc(h, &t.img_tcritic, 2 * tcc::IMG_FWD))) return rc;
            if ((rc = dev_alloc(h, &t.imgt_actor, tcc::IMG_T))) return rc;
            if ((rc = dev_alloc(h, &t.imgt_critic, 2 * tcc::IMG_T))) return rc;
            if ((rc = dev_alloc(h, &t.DH1P, (size_t)tcc::MAX_NCL * tcc::CL * tcc::RB * H1))) return rc;
            if ((rc = dev_alloc(h, &t.GW2PC, (size_t)2 * tcc::MAX_RB * H2 * H1))) return rc;
            if ((rc = dev_alloc(h, &t.GW2PA, (size_t)2 * tcc::MAX_RB * H2 * H1))) return rc;
            if ((rc = dev_alloc(h, &t.PARTC, (size_t)tcc::MAX_RB * tcc::CL * 2 * CriticL::SMALL))) return rc;
            if ((rc = dev_alloc(h, &t.PARTA, (size_t)2 * tcc::MAX_RB * tcc::CL * ActorL::SMALL))) return rc;
            if ((rc = dev_alloc(h, &t.TQX, (size_t)tcc::MAX_RB * 2 * tcc::RB))) return rc;
            if ((rc = dev_alloc(h, &t.RLQX, (size_t)tcc::MAX_RB * tcc::RB))) return rc;
            if ((rc = dev_alloc(h, &t.PIX, (size_t)tcc::MAX_RB * tcc::RB * 4))) return rc;
            if ((rc = dev_alloc(h, &t.RSA, (size_t)tcc::MAX_RB * tcc::RB))) return rc;
            if ((rc = dev_alloc(h, &t.OWNP, (size_t)tcc::MAX_RB * tcc::RB * (H1 + 32)))) return rc;
            if ((rc = dev_alloc(h, &t.XH2S, (size_t)tcc::MAX_RB * tcc::RB * H2))) return rc;
            if ((rc = dev_alloc(h, &t.H1TI, (size_t)tcc::MAX_NCL * tcc::H1T_FLOATS))) return rc;
            if ((rc = dev_alloc(h, &t.APX, (size_t)tcc::MAX_RB * tcc::RB * 4))) return rc;
            if ((rc = dev_alloc(h, &t.XHP, (size_t)2 * tcc::MAX_RB * tcc::RB * H2))) return rc;
            if ((rc = dev_alloc(h, &t.H1I, (size_t)tcc::MAX_NCL * tcc::H1T_FLOATS))) return rc;
            if ((rc = dev_alloc(h, &t.RSUM, (size_t)tcc::MAX_NCL * 8))) return rc;
            if ((rc = dev_alloc(h, &t.LPT, 16))) return rc;
            if ((rc = dev_alloc(h, &t.CNT, 8))) return rc;
            if ((rc = dev_alloc(h, &t.pflags, (size_t)tcc::PF_WORDS * FLAG_STRIDE))) return rc;
            if ((rc = dev_alloc(h, &t.gen_ctr, 1))) return rc;
        }
        if (proto.use_cluster) {
            if ((rc = dev_alloc(h, &c.w2p_actor, (size_t)H2 * cl::WLD))) return rc;
            if ((rc = dev_alloc(h, &c.w2p_tactor, (size_t)H2 * cl::WLD))) return rc;
            if ((rc = dev_alloc(h, &c.w2p_bc, (size_t)H2 * cl::WLD))) return rc;
            if ((rc = dev_alloc(h, &c.w2p_critic, (size_t)2 * H2 * cl::WLD))) return rc;
            if ((rc = dev_alloc(h, &c.w2p_tcritic, (size_t)2 * H2 * cl::WLD))) return rc;
        }
    }
    const size_t K = (size_t)cfg->max_steps * R;
    if ((rc = dev_alloc(h, &h->d_steps, K))) return rc;
    if ((rc = dev_alloc(h, &h->d_ion, K * B))) return rc;
    if ((rc = dev_alloc(h, &h->d_iex, K * B))) return rc;
    if ((rc = dev_alloc(h, &h->d_ibc, K * B))) return rc;
    if ((rc = dev_alloc(h, &h->d_out, K * HIRL_OUT_SCALARS))) return rc;
    if ((rc = dev_alloc(h, &h->d_bar, (size_t)MAX_CTAS * FLAG_STRIDE))) return rc;
    if (R > 1) {
        if ((rc = dev_alloc(h, &h->d_ctxs, (size_t)R))) return rc;
        CU(cudaFuncSetAttribute(hirl_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    }
    *out = h;
    return 0;
}

int hirl_agent_destroy(hirl_agent_t *h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->d_act_s) cudaFree(h->d_act_s);
    if (h->d_act_a) cudaFree(h->d_act_a);
    if (h->nccl_comm && h->dp_own_comm && g_nccl.destroy) g_nccl.destroy(h->nccl_comm);
    if (h->h_pack) cudaFreeHost(h->h_pack);
    for (void *p : h->dp_opened) cudaIpcCloseMemHandle(p);
    for (void *p : h->allocs) cudaFree(p);
    delete h;
    return 0;
}
int hirl_agent_sm_count(hirl_agent_t *h) { return h ? h->grid : 0; }
int hirl_agent_backend(hirl_agent_t *h) {
    if (!h) return fail(HIRL_E_INVALID, "null handle");
    if (h->ctxs.size() > 1) return 3;
    return h->ctxs[0].use_tcc ? 1 : h->ctxs[0].use_cluster ? 2 : 0;
}

// ---- data-parallel mode: communicator ---------------------------------------------------------------------------------
static int dp_enable(hirl_agent *h, void *comm, bool own, int rank, int world, int global_batch) {
    if (h->ctxs.size() != 1) return fail(HIRL_E_UNSUPPORTED, "data-parallel mode is for a single agent per handle");
    if (world < 1 || rank < 0 || rank >= world) return fail(HIRL_E_INVALID, "bad rank / world");
    if (global_batch != h->cfg.batch * world) return fail(HIRL_E_INVALID, "global_batch must be world * cfg.batch (equal shards)");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    if (!h->d_dp_count) { if (int rc = dev_alloc(h, &h->d_dp_count, 1)) return rc; }
    AgentCtx &c = h->ctxs[0];
    // the stage kernel carries the mode (its stages keep their state in global memory, so a call can be cut at the two
    // optimizer boundaries); an agent that was created on the cluster kernels moves over: parameters, moments, replay
    // tables and the stage scratch are shared by all back-ends
    c.use_tcc = 0; c.use_cluster = 0;   // (created with use_tc = tc_g1 = 0: the FP32 tiles, which need no derived weight images)
    c.Bg = global_batch;
    c.dp_count = h->d_dp_count;
    h->nccl_comm = comm; h->dp_own_comm = own; h->dp_rank = rank; h->dp_world = world;
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    return 0;
}

// ---- data-parallel mode, fused variant: all-reduce inside the update kernel over NVLink peer memory -------------------------
// hirl_dp_fused_export allocates this rank's exchange block (common.cuh: DPX_*) and returns its CUDA IPC handle; the caller
// gathers the handles of all ranks by any side channel and hirl_dp_fused_attach maps them.  From then on a call is ONE launch
// (several calls per launch work too): every rank writes its shard's gradient into its block, raises a generation flag and
// sums all ranks' blocks in rank order with P2P loads -- bit-identical on every rank, no NCCL call, no extra launch.
int hirl_dp_fused_export(hirl_agent_t *h, int rank, int world, int global_batch, void *handle64) {
    if (!h || !handle64) return fail(HIRL_E_INVALID, "null argument");
    if (h->ctxs.size() != 1) return fail(HIRL_E_UNSUPPORTED, "data-parallel mode is for a single agent per handle");
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(HIRL_E_INVALID, "bad rank / world (1..8 ranks)");
    if (global_batch != h->cfg.batch * world) return fail(HIRL_E_INVALID, "global_batch must be world * cfg.batch (equal shards)");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    if (!h->d_dp_xchg) {
        if (int rc = dev_alloc(h, &h->d_dp_xchg, DPX_FLOATS)) return rc;
    }
    AgentCtx &c = h->ctxs[0];
    if (!c.dp_gen_ctr) { if (int rc = dev_alloc(h, &c.dp_gen_ctr, 1)) return rc; }
    cudaIpcMemHandle_t ih;
    CU(cudaIpcGetMemHandle(&ih, h->d_dp_xchg));
    static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &ih, 64);
    c.use_tcc = 0; c.use_cluster = 0;          // the stage kernel carries the mode (see dp_enable)
    c.Bg = global_batch;
    c.dp_rank = rank;
    c.dp_world = 0;                            // enabled by hirl_dp_fused_attach
    h->dp_rank = rank;
    return 0;
}
int hirl_dp_fused_attach(hirl_agent_t *h, const void *handles, int world) {
    if (!h || !handles) return fail(HIRL_E_INVALID, "null argument");
    if (!h->d_dp_xchg) return fail(HIRL_E_INVALID, "call hirl_dp_fused_export first");
    if (world < 1 || world > 8) return fail(HIRL_E_INVALID, "bad world");
    CU(cudaSetDevice(h->cfg.device));
    AgentCtx &c = h->ctxs[0];
    for (int r = 0; r < world; ++r) {
        if (r == c.dp_rank) { c.dp_peer[r] = reinterpret_cast<char *>(h->d_dp_xchg); continue; }
        cudaIpcMemHandle_t ih;
        std::memcpy(&ih, static_cast<const char *>(handles) + (size_t)r * 64, 64);
        void *p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
        h->dp_opened.push_back(p);
        c.dp_peer[r] = static_cast<char *>(p);
    }
    c.dp_world = world;
    h->dp_fused = true;
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; h->graph_key[0] = nullptr; }
    return 0;
}
int hirl_nccl_unique_id(void *id128) {
    if (!id128) return fail(HIRL_E_INVALID, "null argument");
    if (int rc = nccl_load()) return rc;
    NC(g_nccl.unique_id(id128));
    return 0;
}
int hirl_allreduce_init(hirl_agent_t *h, const void *id128, int rank, int world, int global_batch) {
    if (!h || !id128) return fail(HIRL_E_INVALID, "null argument");
    if (int rc = nccl_load()) return rc;
    CU(cudaSetDevice(h->cfg.device));
    NcclUid uid;
    std::memcpy(&uid, id128, sizeof(uid));
    void *comm = nullptr;
    NC(g_nccl.init_rank(&comm, world, uid, rank));
    if (int rc = dp_enable(h, comm, true, rank, world, global_batch)) { g_nccl.destroy(comm); return rc; }
    return 0;
}
int hirl_allreduce_attach(hirl_agent_t *h, void *nccl_comm, int rank, int world, int global_batch) {
    if (!h || !nccl_comm) return fail(HIRL_E_INVALID, "null argument");
    if (int rc = nccl_load()) return rc;
    return dp_enable(h, nccl_comm, false, rank, world, global_batch);
}

int hirl_params_import(hirl_agent_t *h, int replica, int net, const float *src, size_t n) {
    if (!h || !src) return fail(HIRL_E_INVALID, "null argument");
    return flat_import(h, replica, net, 0, src, n);
}
int hirl_params_export(hirl_agent_t *h, int replica, int net, float *dst, size_t n) {
    if (!h || !dst) return fail(HIRL_E_INVALID, "null argument");
    return flat_export(h, replica, net, 0, dst, n);
}
int hirl_adam_import(hirl_agent_t *h, int replica, int net, const float *m, const float *v, size_t n) {
    if (!h || !m || !v) return fail(HIRL_E_INVALID, "null argument");
    if (int rc = flat_import(h, replica, net, 2, m, n)) return rc;
    return flat_import(h, replica, net, 3, v, n);
}
int hirl_adam_export(hirl_agent_t *h, int replica, int net, float *m, float *v, size_t n) {
    if (!h || !m || !v) return fail(HIRL_E_INVALID, "null argument");
    if (int rc = flat_export(h, replica, net, 2, m, n)) return rc;
    return flat_export(h, replica, net, 3, v, n);
}
int hirl_grads_export(hirl_agent_t *h, int replica, int net, float *dst, size_t n) {
    if (!h || !dst) return fail(HIRL_E_INVALID, "null argument");
    return flat_export(h, replica, net, 1, dst, n);
}

int hirl_store_rows(hirl_agent_t *h, int replica, int buffer, int first_slot, int n_rows, const float *rows32) {
    if (!h || !rows32 || replica < 0 || replica >= (int)h->ctxs.size()) return fail(HIRL_E_INVALID, "bad handle/replica");
    const int cap = buffer == HIRL_BUF_ONLINE ? h->cfg.online_capacity : buffer == HIRL_BUF_EXPERT ? h->cfg.expert_capacity : -1;
    if (cap < 0) return fail(HIRL_E_INVALID, "bad buffer selector");
    if (n_rows < 0 || n_rows > cap || first_slot < 0 || first_slot >= cap) return fail(HIRL_E_INVALID, "row range outside the ring");
    float *base = buffer == HIRL_BUF_ONLINE ? h->online[replica] : h->expert[replica];
    CU(cudaSetDevice(h->cfg.device));
    const int first = std::min(n_rows, cap - first_slot);
    CU(cudaMemcpyAsync(base + (size_t)first_slot * ROW, rows32, (size_t)first * ROW * 4, cudaMemcpyHostToDevice, h->stream));
    if (n_rows > first) CU(cudaMemcpyAsync(base, rows32 + (size_t)first * ROW, (size_t)(n_rows - first) * ROW * 4, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int hirl_bc_table_set(hirl_agent_t *h, const float *states13, const float *actions4, int n_rows) {
    if (!h || !states13 || !actions4) return fail(HIRL_E_INVALID, "null argument");
    if (n_rows != h->cfg.bc_rows) return fail(HIRL_E_INVALID, "expert table size differs from cfg.bc_rows");
    std::vector<float> tmp((size_t)n_rows * BCW, 0.f);
    for (int i = 0; i < n_rows; ++i) {
        std::memcpy(&tmp[(size_t)i * BCW], states13 + (size_t)i * S, S * 4);
        std::memcpy(&tmp[(size_t)i * BCW + S], actions4 + (size_t)i * A, A * 4);
    }
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpy(h->bctab, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

int hirl_gather(hirl_agent_t *h, int replica, const int32_t *idx_online, int n_online, const int32_t *idx_expert, int n_expert,
                float *out_rows32) {
    if (!h || !out_rows32 || replica < 0 || replica >= (int)h->ctxs.size() || n_online < 0 || n_expert < 0) return fail(HIRL_E_INVALID, "bad argument");
    const int n = n_online + n_expert;
    if (n == 0) return 0;
    for (int i = 0; i < n_online; ++i)
        if (idx_online[i] < 0 || idx_online[i] >= h->cfg.online_capacity) return fail(HIRL_E_INVALID, "online index out of range");
    for (int i = 0; i < n_expert; ++i)
        if (idx_expert[i] < 0 || idx_expert[i] >= h->cfg.expert_capacity) return fail(HIRL_E_INVALID, "expert index out of range");
    CU(cudaSetDevice(h->cfg.device));
    int32_t *d_idx = nullptr;
    float *d_out = nullptr;
    CU(cudaMalloc(&d_idx, (size_t)n * 4));
    CU(cudaMalloc(&d_out, (size_t)n * ROW * 4));
    if (n_online) CU(cudaMemcpyAsync(d_idx, idx_online, (size_t)n_online * 4, cudaMemcpyHostToDevice, h->stream));
    if (n_expert) CU(cudaMemcpyAsync(d_idx + n_online, idx_expert, (size_t)n_expert * 4, cudaMemcpyHostToDevice, h->stream));
    gather_kernel<<<std::min((n + 7) / 8, 4 * h->sm_count), 256, 0, h->stream>>>(h->online[replica], h->expert[replica], d_idx, n_online, d_idx + n_online, n_expert, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_rows32, d_out, (size_t)n * ROW * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d_idx);
    cudaFree(d_out);
    return 0;
}

int hirl_gather_bc(hirl_agent_t *h, const int32_t *idx_bc, int n, float *out_states13, float *out_actions4) {
    if (!h || !idx_bc || !out_states13 || !out_actions4 || n < 0) return fail(HIRL_E_INVALID, "bad argument");
    if (n == 0) return 0;
    for (int i = 0; i < n; ++i)
        if (idx_bc[i] < 0 || idx_bc[i] >= h->cfg.bc_rows) return fail(HIRL_E_INVALID, "expert-table index out of range");
    CU(cudaSetDevice(h->cfg.device));
    int32_t *d_idx = nullptr;
    float *d_s = nullptr, *d_a = nullptr;
    CU(cudaMalloc(&d_idx, (size_t)n * 4));
    CU(cudaMalloc(&d_s, (size_t)n * S * 4));
    CU(cudaMalloc(&d_a, (size_t)n * A * 4));
    CU(cudaMemcpyAsync(d_idx, idx_bc, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    gather_bc_kernel<<<std::min((n + 7) / 8, 4 * h->sm_count), 256, 0, h->stream>>>(h->bctab, d_idx, n, d_s, d_a);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_states13, d_s, (size_t)n * S * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(out_actions4, d_a, (size_t)n * A * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(d_idx); cudaFree(d_s); cudaFree(d_a);
    return 0;
}

// steps/idx/out are device arrays with a leading [replicas] dimension of stride max(k_steps)

// ---- data-parallel large-batch mode ----------------------------------------------------------------------------------
// One process per GPU, every process a full replica of the agent and of its replay tables; a call's batch of Bg rows is cut
// into world shards of B rows (the host mirror slices the index streams it drew once, hirl4ucav_b200/parallel.py).  The
// stage kernel runs the call as launches over stage ranges and the per-shard gradients meet in NCCL all-reduces over
// NVLink exactly where the reference calls optimizer.step() (HIRL.py:286-288 critic, :323-325 actor):
//   [R0 .. G4 phase 1]  critic forward / backward, shard gradient -> g_critic      all-reduce g_critic (276 482 + pad), loss
//   [G4 phase 2]        critic Adam on the summed gradient (identical on every rank)
//   [R4 .. G6]          Q1(s, pi), Q1(s, a_bc), shard count of soft_Q > rl_Q        all-reduce the count (HIRL.py:303 is a batch mean)
//   [R6 .. G8 phase 1]  actor backward with the global weight w -> g_actor          all-reduce g_actor (138 756 + pad), 4 loss scalars
//   [G8 phase 2]        actor Adam, Polyak, returned scalars
// Loss means are taken over Bg in every shard, so sums over ranks are the global means.  libnccl is resolved at run time
// (the copy torch already loaded, else libnccl.so.2): the library has no link-time dependency on it.
static int launch_update_dp(hirl_agent *h, LaunchCtx &L, int k_steps, cudaStream_t stream) {
    if ((int)h->dp_do_actor.size() < k_steps) return fail(HIRL_E_INVALID, "data-parallel mode needs host step records (hirl_learn / hirl_learn_async)");
    AgentCtx &c = L.agent;
    int launches = 0;
    auto seg = [&](int step, int first, int last, int phase) -> int {
        L.step0 = step; L.k_steps = 1; L.first_stage = first; L.last_stage = last; L.adam_phase = phase; L.stop_after = 0;
        void *args[] = {&L};
        CU(cudaLaunchCooperativeKernel((void *)hirl_update_kernel, dim3(h->grid), dim3(THREADS), args, SMEM_BYTES, stream));
        ++launches;
        return 0;
    };
    for (int step = 0; step < k_steps; ++step) {
        const bool do_actor = h->dp_do_actor[step] != 0;
        if (h->dp_mode[step] != 0) return fail(HIRL_E_UNSUPPORTED, "data-parallel mode implements the HIRL step (mode 0)");
        if (int rc = seg(step, ST_R0, ST_G4, 1)) return rc;
        NC(g_nccl.group_start());
        NC(g_nccl.allreduce(c.g_critic, c.g_critic, (size_t)CRITIC_SIZE, NCCL_FLOAT32, NCCL_SUM, h->nccl_comm, stream));
        NC(g_nccl.allreduce(c.carry, c.carry, 1, NCCL_FLOAT32, NCCL_SUM, h->nccl_comm, stream));
        NC(g_nccl.group_end());
        if (int rc = seg(step, ST_G4, ST_G4, 2)) return rc;
        if (!do_actor) continue;
        if (int rc = seg(step, ST_R4, ST_G6, 0)) return rc;
        NC(g_nccl.allreduce(h->d_dp_count, h->d_dp_count, 1, NCCL_FLOAT32, NCCL_SUM, h->nccl_comm, stream));
        if (int rc = seg(step, ST_R6, ST_G8, 1)) return rc;
        NC(g_nccl.group_start());
        NC(g_nccl.allreduce(c.g_actor, c.g_actor, (size_t)ACTOR_SIZE, NCCL_FLOAT32, NCCL_SUM, h->nccl_comm, stream));
        NC(g_nccl.allreduce(c.carry + 1, c.carry + 1, 4, NCCL_FLOAT32, NCCL_SUM, h->nccl_comm, stream));
        NC(g_nccl.group_end());
        if (int rc = seg(step, ST_G8, ST_G8, 2)) return rc;
    }
    h->last_launches = launches;
    return 0;
}

static int launch_update(hirl_agent *h, int k_steps, const hirl_step_t *steps_dev, const int32_t *ion, const int32_t *iex, const int32_t *ibc,
                         float *out_dev, cudaStream_t stream) {
    const int R = (int)h->ctxs.size();
    const size_t B = h->cfg.batch;
    // Barrier flags are never reset: every CTA of a team leaves a launch with the same generation in its flag and
    // resumes from it (no memset per launch, no per-launch kernel argument; the comparison in team_sync is wrap-safe).
    if (R == 1 && h->ctxs[0].use_tcc) {
        tcc::TccLaunch L;
        L.agent = h->ctxs[0];
        AgentCtx &c = L.agent;
        c.steps = steps_dev; c.idx_online = ion; c.idx_expert = iex; c.idx_bc = ibc; c.out_scalars = out_dev;
        L.bar_flags = h->d_bar; L.k_steps = k_steps; L.ncl = h->tcc_ncl;
        L.prof = h->prof_on ? h->d_prof : nullptr;
        { const char *e = getenv("HIRL_TCC_JITTER"); L.jitter = e ? (unsigned)atoi(e) : 0u; }
        { const char *e = getenv("HIRL_TCC_JITTER_NS"); L.jitter_ns = e ? (unsigned)atoi(e) : 0u; }
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(h->tcc_ncl * tcc::CL); lc.blockDim = dim3(tcc::TTHREADS);
        lc.dynamicSmemBytes = tcc::SMEM_BYTES;
        lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;     // clusters spin on each other's flags: they must be co-resident
        lc.attrs = at; lc.numAttrs = h->coop_ok ? 1 : 0;
        void *args[] = {&L};
        cudaError_t e = cudaLaunchKernelExC(&lc, (const void *)tcc::hirl_tcc_kernel, args);
        if (e != cudaSuccess && h->coop_ok) {
            cudaGetLastError();
            h->coop_ok = false;
            lc.numAttrs = 0;
            e = cudaLaunchKernelExC(&lc, (const void *)tcc::hirl_tcc_kernel, args);
        }
        if (e != cudaSuccess) return fail(HIRL_E_CUDA, std::string("tcc kernel launch: ") + cudaGetErrorString(e));
#ifndef HIRL_NO_CLUSTER
    } else if (R == 1 && h->ctxs[0].use_cluster) {
        cl::ClusterLaunch L;
        L.agent = h->ctxs[0];
        AgentCtx &c = L.agent;
        c.steps = steps_dev; c.idx_online = ion; c.idx_expert = iex; c.idx_bc = ibc; c.out_scalars = out_dev;
        L.bar_flags = h->d_bar; L.k_steps = k_steps; L.ncl = h->ncl;
        L.prof = h->prof_on ? h->d_prof : nullptr;
        const bool r9 = c.cluster_ri == 9;
        const void *fn = r9 ? (const void *)cl::hirl_cluster_kernel<9> : (const void *)cl::hirl_cluster_kernel<8>;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(h->ncl * cl::CL); lc.blockDim = dim3(cl::CTHREADS);
        lc.dynamicSmemBytes = r9 ? cl::Ker<9>::SMEM_BYTES : cl::Ker<8>::SMEM_BYTES;
        lc.stream = stream;
        // cooperative: the clusters of one launch spin on each other's flags, so they must be co-resident (two agents on
        // two streams must not interleave half-resident grids); the cluster shape is part of the kernel (__cluster_dims__)
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        lc.attrs = at; lc.numAttrs = h->coop_ok ? 1 : 0;
        void *args[] = {&L};
        cudaError_t e = cudaLaunchKernelExC(&lc, fn, args);
        if (e != cudaSuccess && h->coop_ok) {   // a driver that refuses cooperative + cluster launches: plain launch (grid <= co-resident clusters)
            cudaGetLastError();
            h->coop_ok = false;
            lc.numAttrs = 0;
            e = cudaLaunchKernelExC(&lc, fn, args);
        }
        if (e != cudaSuccess) return fail(HIRL_E_CUDA, std::string("cluster kernel launch: ") + cudaGetErrorString(e));
#endif
    } else if (R == 1) {
        LaunchCtx L;
        L.agent = h->ctxs[0];
        AgentCtx &c = L.agent;
        c.steps = steps_dev; c.idx_online = ion; c.idx_expert = iex; c.idx_bc = ibc; c.out_scalars = out_dev;
        L.bar_flags = h->d_bar; L.k_steps = k_steps; L.stop_after = h->stop_after;
        L.prof = h->prof_on ? h->d_prof : nullptr;
        L.step0 = 0; L.first_stage = 1; L.last_stage = ST_END; L.adam_phase = 0;
        if (h->dp_world > 0) return launch_update_dp(h, L, k_steps, stream);
        if (h->dp_fused) L.adam_phase = 4;      // whole calls in one launch, all-reduces inside the kernel
        void *args[] = {&L};
        CU(cudaLaunchCooperativeKernel((void *)hirl_update_kernel, dim3(h->grid), dim3(THREADS), args, SMEM_BYTES, stream));
    } else {
        std::vector<AgentCtx> tmp(h->ctxs);
        for (int r = 0; r < R; ++r) {
            AgentCtx &c = tmp[r];
            const size_t o = (size_t)r * k_steps;
            c.steps = steps_dev + o; c.idx_online = ion + o * B; c.idx_expert = iex + o * B; c.idx_bc = ibc + o * B;
            c.out_scalars = out_dev + o * HIRL_OUT_SCALARS;
        }
        CU(cudaMemcpyAsync(h->d_ctxs, tmp.data(), sizeof(AgentCtx) * R, cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));  // tmp is pageable host memory
        EnsembleLaunch L;
        L.agents = h->d_ctxs; L.bar_flags = h->d_bar; L.replicas = R; L.n_teams = h->n_teams; L.team_size = h->team_size; L.k_steps = k_steps;
        void *args[] = {&L};
        CU(cudaLaunchCooperativeKernel((void *)hirl_ensemble_kernel, dim3(h->n_teams * h->team_size), dim3(THREADS), args, SMEM_BYTES, stream));
    }
    h->last_launches = 1;
    return 0;
}

static int check_steps(hirl_agent *h, int k_steps, const hirl_step_t *steps) {
    if (k_steps < 1 || k_steps > h->cfg.max_steps) return fail(HIRL_E_INVALID, "k_steps outside [1, cfg.max_steps]");
    for (int k = 0; k < k_steps * (int)h->ctxs.size(); ++k) {
        const hirl_step_t &s = steps[k];
        if (s.mode != 2 && (s.n_online < 0 || s.n_expert < 0 || s.n_online + s.n_expert != h->cfg.batch))
            return fail(HIRL_E_INVALID, "n_online + n_expert must equal the batch size");
        if (s.mode < 0 || s.mode > 2) return fail(HIRL_E_INVALID, "bad step mode");
        if (s.mode != 0 && h->dp_fused) return fail(HIRL_E_UNSUPPORTED, "data-parallel mode implements the HIRL step (mode 0)");
        if (s.mode != 0 && h->ctxs[0].use_tcc) return fail(HIRL_E_INVALID, "TD3 / BC step modes run on the stage kernel (create the agent with HIRL_TCC=0)");
        if (s.mode != 1 && s.do_actor && h->cfg.bc_rows < h->cfg.batch) return fail(HIRL_E_INVALID, "expert table smaller than the batch (np.random.choice without replacement)");
    }
    return 0;
}

int hirl_learn(hirl_agent_t *h, int k_steps, const hirl_step_t *steps, const int32_t *idx_online, const int32_t *idx_expert,
               const int32_t *idx_bc, float *out_scalars) {
    if (!h || !steps || !idx_online || !out_scalars) return fail(HIRL_E_INVALID, "null argument");
    if (int rc = check_steps(h, k_steps, steps)) return rc;
    if (h->dp_world > 0) {   // the data-parallel launcher cuts every call into stage ranges on the host: it needs the control words
        h->dp_do_actor.assign(k_steps, 0); h->dp_mode.assign(k_steps, 0);
        for (int k = 0; k < k_steps; ++k) { h->dp_do_actor[k] = steps[k].do_actor; h->dp_mode[k] = steps[k].mode; }
    }
    const size_t B = h->cfg.batch, K = (size_t)k_steps * h->ctxs.size();
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h->d_steps, steps, K * sizeof(hirl_step_t), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_ion, idx_online, K * B * 4, cudaMemcpyHostToDevice, h->stream));
    if (idx_expert) CU(cudaMemcpyAsync(h->d_iex, idx_expert, K * B * 4, cudaMemcpyHostToDevice, h->stream));
    if (idx_bc) CU(cudaMemcpyAsync(h->d_ibc, idx_bc, K * B * 4, cudaMemcpyHostToDevice, h->stream));
    if (int rc = launch_update(h, k_steps, h->d_steps, h->d_ion, h->d_iex, h->d_ibc, h->d_out, h->stream)) return rc;
    CU(cudaMemcpyAsync(out_scalars, h->d_out, K * HIRL_OUT_SCALARS * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// true when p is pinned host memory the GPU can address directly (UVA: cudaHostAlloc / torch pin_memory)
static bool device_addressable_host(const void *p, const void **dev) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dev = a.devicePointer;
    return true;
}

int hirl_learn_async(hirl_agent_t *h, int k_steps, const hirl_step_t *steps, const int32_t *idx_online, const int32_t *idx_expert,
                     const int32_t *idx_bc, float *out_scalars) {
    if (!h || !steps || !idx_online || !out_scalars) return fail(HIRL_E_INVALID, "null argument");
    if (int rc = check_steps(h, k_steps, steps)) return rc;
    if (h->dp_world > 0) {   // the data-parallel launcher cuts every call into stage ranges on the host: it needs the control words
        h->dp_do_actor.assign(k_steps, 0); h->dp_mode.assign(k_steps, 0);
        for (int k = 0; k < k_steps; ++k) { h->dp_do_actor[k] = steps[k].do_actor; h->dp_mode[k] = steps[k].mode; }
    }
    const size_t B = h->cfg.batch, K = (size_t)k_steps * h->ctxs.size();
    CU(cudaSetDevice(h->cfg.device));
    // Optional zero-copy (off by default): with every stream in pinned host memory the kernel can read the step
    // records and indices straight over PCIe and write its 8 scalars straight back, which saves five
    // cudaMemcpyAsync calls per learn().  Measured on B200: the PCIe round trips on the kernel's critical path cost
    // more than the copies (e2e 5 075 vs 5 257 updates/s, learn_many 6 340 vs 7 076), so the staged copies stay.
    const void *d_st, *d_on, *d_ex = nullptr, *d_bc = nullptr, *d_out;
    if (h->zero_copy && h->dp_world == 0 && device_addressable_host(steps, &d_st) && device_addressable_host(idx_online, &d_on) &&
        (!idx_expert || device_addressable_host(idx_expert, &d_ex)) && (!idx_bc || device_addressable_host(idx_bc, &d_bc)) &&
        device_addressable_host(out_scalars, &d_out)) {
        return launch_update(h, k_steps, (const hirl_step_t *)d_st, (const int32_t *)d_on, d_ex ? (const int32_t *)d_ex : h->d_iex,
                             d_bc ? (const int32_t *)d_bc : h->d_ibc, (float *)d_out, h->stream);
    }
    auto enqueue = [&]() -> int {
        CU(cudaMemcpyAsync(h->d_steps, steps, K * sizeof(hirl_step_t), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->d_ion, idx_online, K * B * 4, cudaMemcpyHostToDevice, h->stream));
        if (idx_expert) CU(cudaMemcpyAsync(h->d_iex, idx_expert, K * B * 4, cudaMemcpyHostToDevice, h->stream));
        if (idx_bc) CU(cudaMemcpyAsync(h->d_ibc, idx_bc, K * B * 4, cudaMemcpyHostToDevice, h->stream));
        if (int rc = launch_update(h, k_steps, h->d_steps, h->d_ion, h->d_iex, h->d_ibc, h->d_out, h->stream)) return rc;
        CU(cudaMemcpyAsync(out_scalars, h->d_out, K * HIRL_OUT_SCALARS * 4, cudaMemcpyDeviceToHost, h->stream));
        return 0;
    };
    const bool graphable = h->use_graph && k_steps == 1 && h->ctxs.size() == 1 && !h->prof_on && h->stop_after == 0 && h->dp_world == 0;
    if (graphable) {
        // One call = [packed H2D copy, cooperative launch, D2H of 8 scalars], captured once into a CUDA graph.  The caller's four
        // staging pieces (64 B + 3 x B x 4 B) are gathered by the CPU into a library-owned pinned block first, so the graph
        // has one copy node instead of four and does not depend on the caller's staging addresses.  The block is reused: the
        // stream must be idle (it is, for the per-call API: learn() waits for its result); otherwise plain calls.
        const size_t rowb = B * 4, need = sizeof(hirl_step_t) + 3 * rowb;
        if (!h->h_pack) {
            if (cudaHostAlloc((void **)&h->h_pack, need, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); h->h_pack = nullptr; }
            else if (dev_alloc(h, &h->d_pack, need)) { cudaFreeHost(h->h_pack); h->h_pack = nullptr; }
            h->pack_bytes = need;
        }
        if (!h->h_pack || cudaStreamQuery(h->stream) != cudaSuccess) { cudaGetLastError(); return enqueue(); }
        std::memcpy(h->h_pack, steps, sizeof(hirl_step_t));
        std::memcpy(h->h_pack + sizeof(hirl_step_t), idx_online, rowb);
        if (idx_expert) std::memcpy(h->h_pack + sizeof(hirl_step_t) + rowb, idx_expert, rowb);
        if (idx_bc) std::memcpy(h->h_pack + sizeof(hirl_step_t) + 2 * rowb, idx_bc, rowb);
        auto enqueue_packed = [&]() -> int {
            CU(cudaMemcpyAsync(h->d_pack, h->h_pack, need, cudaMemcpyHostToDevice, h->stream));
            const char *d = h->d_pack;
            if (int rc = launch_update(h, 1, (const hirl_step_t *)d, (const int32_t *)(d + sizeof(hirl_step_t)),
                                       (const int32_t *)(d + sizeof(hirl_step_t) + rowb), (const int32_t *)(d + sizeof(hirl_step_t) + 2 * rowb),
                                       h->d_out, h->stream)) return rc;
            CU(cudaMemcpyAsync(out_scalars, h->d_out, HIRL_OUT_SCALARS * 4, cudaMemcpyDeviceToHost, h->stream));
            return 0;
        };
        const void *key[5] = {out_scalars, nullptr, nullptr, nullptr, nullptr};
        if (!h->graph_exec && !h->graph_key[0]) {   // built once, for the first result buffer seen
            cudaGraph_t g = nullptr;
            bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                ok = enqueue_packed() == 0;
                ok = (cudaStreamEndCapture(h->stream, &g) == cudaSuccess) && ok && g;
            }
            if (ok) ok = cudaGraphInstantiate(&h->graph_exec, g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
            if (!ok) {   // e.g. a driver that refuses cooperative launches inside a capture: plain calls from now on
                cudaGetLastError();
                h->graph_exec = nullptr;
                h->use_graph = false;
            }
            std::memcpy(h->graph_key, key, sizeof(key));
        }
        if (h->graph_exec && std::memcmp(key, h->graph_key, sizeof(key)) != 0) return enqueue_packed();   // another result buffer: plain calls
        if (h->graph_exec) {
            CU(cudaGraphLaunch(h->graph_exec, h->stream));
            h->last_launches = 1;
            return 0;
        }
        return enqueue_packed();
    }
    return enqueue();
}
int hirl_learn_wait(hirl_agent_t *h) {
    if (!h) return fail(HIRL_E_INVALID, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int hirl_learn_dev(hirl_agent_t *h, int k_steps, const hirl_step_t *steps_dev, const int32_t *idx_online_dev, const int32_t *idx_expert_dev,
                   const int32_t *idx_bc_dev, float *out_scalars_dev, void *stream) {
    if (!h || !steps_dev || !idx_online_dev || !out_scalars_dev) return fail(HIRL_E_INVALID, "null argument");
    if (k_steps < 1) return fail(HIRL_E_INVALID, "k_steps must be >= 1");
    if (h->dp_world > 0) return fail(HIRL_E_UNSUPPORTED, "data-parallel agents take host step records (hirl_learn / hirl_learn_async)");
    CU(cudaSetDevice(h->cfg.device));
    return launch_update(h, k_steps, steps_dev, idx_online_dev, idx_expert_dev ? idx_expert_dev : h->d_iex, idx_bc_dev ? idx_bc_dev : h->d_ibc,
                         out_scalars_dev, (cudaStream_t)stream);
}
int hirl_last_launch_count(hirl_agent_t *h) { return h ? h->last_launches : 0; }

int hirl_act(hirl_agent_t *h, int replica, int net, const float *states13, int n, float *actions4) {
    if (!h || !states13 || !actions4 || n < 0) return fail(HIRL_E_INVALID, "bad argument");
    if (is_critic(net)) return fail(HIRL_E_INVALID, "hirl_act needs an actor network");
    if (n == 0) return 0;
    float *w;
    if (int rc = net_ptr(h, replica, net, 0, &w)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    if (n > h->act_cap) {   // staging grows on demand and is kept: chooseAction runs once per env step (no malloc/free per call)
        const int cap = std::max(n, std::max(64, 2 * h->act_cap));
        float *ns = nullptr, *na = nullptr;
        CU(cudaMalloc(&ns, (size_t)cap * S * 4));
        CU(cudaMalloc(&na, (size_t)cap * A * 4));
        CU(cudaStreamSynchronize(h->stream));
        if (h->d_act_s) cudaFree(h->d_act_s);
        if (h->d_act_a) cudaFree(h->d_act_a);
        h->d_act_s = ns; h->d_act_a = na; h->act_cap = cap;
    }
    CU(cudaMemcpyAsync(h->d_act_s, states13, (size_t)n * S * 4, cudaMemcpyHostToDevice, h->stream));
    act_kernel<<<std::min(n, 8 * h->sm_count), THREADS, 0, h->stream>>>(w, h->d_act_s, n, h->d_act_a, h->cfg.activation == HIRL_ACT_LEAKY_RELU ? 0.01f : 0.f);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(actions4, h->d_act_a, (size_t)n * A * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int hirl_act_dev(hirl_agent_t *h, int replica, int net, const float *states13_dev, int n, float *actions4_dev,
                 const float *noise4_dev, int noise_rows, void *stream) {
    if (!h || !states13_dev || !actions4_dev || n < 0) return fail(HIRL_E_INVALID, "bad argument");
    if (is_critic(net)) return fail(HIRL_E_INVALID, "hirl_act_dev needs an actor network");
    if (noise4_dev && noise_rows != 1 && noise_rows != n) return fail(HIRL_E_INVALID, "noise_rows must be 1 or n");
    if (n == 0) return 0;
    float *w;
    if (int rc = net_ptr(h, replica, net, 0, &w)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    act_kernel<<<std::min(n, 8 * h->sm_count), THREADS, 0, (cudaStream_t)stream>>>(
        w, states13_dev, n, actions4_dev, h->cfg.activation == HIRL_ACT_LEAKY_RELU ? 0.01f : 0.f, noise4_dev,
        noise4_dev ? (noise_rows == 1 ? 0 : A) : 0);
    CU(cudaGetLastError());
    return 0;
}

int hirl_store_transitions_dev(hirl_agent_t *h, int replica, int buffer, int first_slot, int n, const float *s13_dev,
                               const float *a4_dev, const float *s2_13_dev, const float *reward_dev,
                               const int32_t *flags_dev, void *stream) {
    if (!h || replica < 0 || replica >= (int)h->ctxs.size()) return fail(HIRL_E_INVALID, "bad handle/replica");
    if (!s13_dev || !a4_dev || !s2_13_dev || !reward_dev || !flags_dev) return fail(HIRL_E_INVALID, "null argument");
    const int cap = buffer == HIRL_BUF_ONLINE ? h->cfg.online_capacity : buffer == HIRL_BUF_EXPERT ? h->cfg.expert_capacity : -1;
    if (cap < 0) return fail(HIRL_E_INVALID, "bad buffer selector");
    if (n < 0 || n > cap || first_slot < 0 || first_slot >= cap) return fail(HIRL_E_INVALID, "row range outside the ring");
    if (n == 0) return 0;
    float *base = buffer == HIRL_BUF_ONLINE ? h->online[replica] : h->expert[replica];
    CU(cudaSetDevice(h->cfg.device));
    store_transitions_kernel<<<std::min((n + 7) / 8, 8 * h->sm_count), 256, 0, (cudaStream_t)stream>>>(
        base, cap, first_slot, n, s13_dev, a4_dev, s2_13_dev, reward_dev, flags_dev);
    CU(cudaGetLastError());
    return 0;
}

int hirl_agent_stream(hirl_agent_t *h, void **stream_out) {
    if (!h || !stream_out) return fail(HIRL_E_INVALID, "null argument");
    *stream_out = (void *)h->stream;
    return 0;
}

static bool choice_args_ok(const hirl_mt19937_t *st, int32_t population, int32_t size, const int32_t *out) {
    return st && out && population >= 1 && size >= 0 && size <= population && st->pos >= 0 && st->pos <= 624;
}
static void choice_noreplace_impl(hirl_mt19937_t *st, int32_t population, int32_t size, int32_t *out) {
    hirl_host::MT19937 g{st->key, st->pos};
    static thread_local std::vector<int32_t> perm;
    perm.resize((size_t)population);
    int32_t *p = perm.data();
    for (int32_t i = 0; i < population; ++i) p[i] = i;
    // mtrand _shuffle_raw: for i = n-1 .. 1: j = random_interval(i) (masked rejection on one 32-bit output per try),
    // swap(i, j).  Written branch-free: a rejected output swaps p[i] with itself and does not advance i (the
    // data-dependent rejection branch was half the run time); i runs through segments (2^b-1 .. 2^(b-1)] in which
    // the mask is constant, which keeps the loop-carried chain at and/compare/subtract; outputs are tempered a
    // block at a time and the unconsumed tail of the last block is handed back to the generator.
    uint32_t blk[624];
    int32_t i = population - 1;
    int q = 0, avail = 0;
    while (i >= 1) {
        const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)i);
        const int32_t lo = (int32_t)(mask >> 1);
        while (i > lo) {
            if (q == avail) {
                if (g.pos >= 624) g.regenerate();
                avail = 624 - g.pos;
                const uint32_t *kp = g.key + g.pos;
                for (int t = 0; t < avail; ++t) blk[t] = hirl_host::MT19937::temper(kp[t]);
                g.pos = 624;
                q = 0;
            }
            int qq = q;
            int32_t ii = i;
            for (; qq < avail && ii > lo; ++qq) {
                const uint32_t v = blk[qq] & mask;
                const bool ok = v <= (uint32_t)ii;
                const uint32_t j = ok ? v : (uint32_t)ii;
                const int32_t t = p[j];
                p[j] = p[ii];
                p[ii] = t;
                ii -= ok;
            }
            q = qq;
            i = ii;
        }
    }
    if (avail) g.pos = 624 - (avail - q);   // hand the unconsumed outputs back (no block drawn: position untouched)
    for (int32_t i = 0; i < size; ++i) out[i] = perm[i];
    st->pos = g.pos;
}
int hirl_host_choice_noreplace(hirl_mt19937_t *st, int32_t population, int32_t size, int32_t *out) {
    if (!choice_args_ok(st, population, size, out)) return fail(HIRL_E_INVALID, "bad argument");
    choice_noreplace_impl(st, population, size, out);
    return 0;
}

// The same draw on a library-owned worker thread: np.random.choice(20000, 256, replace=False) is a full 20 000-element shuffle
// (~60 us), longer than everything else the host does for one learn() call.  The caller hands over a private COPY of the
// generator state (speculative pre-draw, utils/fast_rng.Predraw) and collects the result at the next call; one job at a time.
namespace {
struct HostWorker {
    std::mutex m;
    std::condition_variable cv;
    bool has_job = false, done = true;
    hirl_mt19937_t *st = nullptr;
    int32_t n = 0, k = 0, count = 1;
    int32_t *out = nullptr;
};
HostWorker *g_worker = nullptr;   // leaked on purpose: the detached thread may outlive static destructors
void host_worker_main(HostWorker *w) {
    for (;;) {
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [w] { return w->has_job; });
        hirl_mt19937_t *st = w->st;
        const int32_t n = w->n, k = w->k, count = w->count;
        int32_t *out = w->out;
        w->has_job = false;
        lk.unlock();
        for (int32_t i = 0; i < count; ++i) choice_noreplace_impl(st, n, k, out + (size_t)i * k);
        lk.lock();
        w->done = true;
        w->cv.notify_all();
    }
}
}  // namespace
int hirl_host_choice_many_async(hirl_mt19937_t *st, int32_t population, int32_t size, int32_t count, int32_t *out) {
    if (!choice_args_ok(st, population, size, out) || count < 1) return fail(HIRL_E_INVALID, "bad argument");
    if (!g_worker) {
        g_worker = new HostWorker;
        std::thread(host_worker_main, g_worker).detach();
    }
    HostWorker *w = g_worker;
    std::lock_guard<std::mutex> lk(w->m);
    if (!w->done) return fail(HIRL_E_UNSUPPORTED, "a host draw is already pending (one at a time)");
    w->st = st; w->n = population; w->k = size; w->count = count; w->out = out;
    w->done = false; w->has_job = true;
    w->cv.notify_all();
    return 0;
}
int hirl_host_choice_noreplace_async(hirl_mt19937_t *st, int32_t population, int32_t size, int32_t *out) {
    return hirl_host_choice_many_async(st, population, size, 1, out);
}
int hirl_host_async_wait(void) {
    HostWorker *w = g_worker;
    if (!w) return 0;
    std::unique_lock<std::mutex> lk(w->m);
    w->cv.wait(lk, [w] { return w->done; });
    return 0;
}

int hirl_host_sample_indices(hirl_mt19937_t *st, int32_t n, int32_t k, int32_t *out) {
    if (!st || !out || n < 0 || k < 0 || k > n || st->pos < 0 || st->pos > 624) return fail(HIRL_E_INVALID, "bad argument");
    if (k == 0) return 0;
    hirl_host::MT19937 g{st->key, st->pos};
    auto randbelow = [&g](uint32_t m) {   // Random._randbelow_with_getrandbits: getrandbits(m.bit_length()), redrawn while >= m
        int bits = 0;
        for (uint32_t t = m; t; t >>= 1) ++bits;
        uint32_t r;
        do { r = g.next() >> (32 - bits); } while (r >= m);
        return r;
    };
    // Random.sample: setsize = 21 (+ 4 ** ceil(log(3k, 4)) for k > 5) decides between the pool and the set algorithm
    long long setsize = 21;
    if (k > 5) {
        long long p = 1;
        while (p < 3LL * k) p *= 4;       // 4 ** ceil(log_4(3k)) for integer 3k
        setsize += p;
    }
    if (n <= setsize) {
        static thread_local std::vector<int32_t> pool;
        pool.resize((size_t)n);
        for (int32_t i = 0; i < n; ++i) pool[i] = i;
        for (int32_t i = 0; i < k; ++i) {
            const uint32_t j = randbelow((uint32_t)(n - i));
            out[i] = pool[j];
            pool[j] = pool[n - i - 1];
        }
    } else {
        // open-addressing set of the selected positions (capacity >= 2k, power of two)
        static thread_local std::vector<int32_t> tab;
        size_t cap = 16;
        while (cap < (size_t)k * 2) cap <<= 1;
        tab.assign(cap, -1);
        for (int32_t i = 0; i < k; ++i) {
            for (;;) {
                const int32_t j = (int32_t)randbelow((uint32_t)n);
                size_t hpos = ((uint32_t)j * 2654435761u) & (cap - 1);
                bool dup = false;
                while (tab[hpos] != -1) {
                    if (tab[hpos] == j) { dup = true; break; }
                    hpos = (hpos + 1) & (cap - 1);
                }
                if (dup) continue;
                tab[hpos] = j;
                out[i] = j;
                break;
            }
        }
    }
    st->pos = g.pos;
    return 0;
}

// Every host index draw of a sweep of `k_calls` learn() calls for `n_jobs` independent generator pairs (ensemble replicas),
// spread over worker threads: job j runs, per call and in the reference's order, random.sample(range(online_len), n_online[k])
// [, random.sample(range(expert_len), n_expert[k])] on ITS py state and np.random.choice(bc_population, batch, replace=False)
// on ITS np state.  Generators of different jobs are independent, so the result does not depend on the thread count.
int hirl_host_draw_many(const hirl_draw_job_t *jobs, int32_t n_jobs, int32_t k_calls, const int32_t *n_online, const int32_t *n_expert,
                        int32_t bc_population, int32_t batch, int32_t *idx_online, int32_t *idx_expert, int32_t *idx_bc, int32_t n_threads) {
    if (!jobs || !n_online || !n_expert || !idx_online || !idx_expert || n_jobs < 0 || k_calls < 1 || batch < 1)
        return fail(HIRL_E_INVALID, "bad argument");
    for (int32_t j = 0; j < n_jobs; ++j) {
        const hirl_draw_job_t &jb = jobs[j];
        if (!jb.py || jb.py->pos < 0 || jb.py->pos > 624 || (idx_bc && (!jb.np || jb.np->pos < 0 || jb.np->pos > 624))) return fail(HIRL_E_INVALID, "bad generator state");
        for (int32_t k = 0; k < k_calls; ++k)
            if (n_online[k] < 0 || n_expert[k] < 0 || n_online[k] + n_expert[k] != batch || n_online[k] > jb.online_len || n_expert[k] > jb.expert_len)
                return fail(HIRL_E_INVALID, "sample larger than population or batch split mismatch");
    }
    if (idx_bc && (bc_population < batch)) return fail(HIRL_E_INVALID, "expert table smaller than the batch");
    const int nt = std::max(1, std::min<int>(n_threads, n_jobs));
    auto work = [&](int t) {
        for (int32_t j = t; j < n_jobs; j += nt) {
            const hirl_draw_job_t &jb = jobs[j];
            for (int32_t k = 0; k < k_calls; ++k) {
                const size_t row = ((size_t)j * k_calls + k) * batch;
                hirl_host_sample_indices(jb.py, jb.online_len, n_online[k], idx_online + row);
                if (n_expert[k]) hirl_host_sample_indices(jb.py, jb.expert_len, n_expert[k], idx_expert + row);
                if (idx_bc) choice_noreplace_impl(jb.np, bc_population, batch, idx_bc + row);
            }
        }
    };
    if (nt == 1) { work(0); return 0; }
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    return 0;
}

int hirl_debug_stop_after(hirl_agent_t *h, int stage) {
    if (!h || stage < 0 || stage >= ST_END) return fail(HIRL_E_INVALID, "bad stage");
    h->stop_after = stage;
    return 0;
}
int hirl_debug_profile(hirl_agent_t *h, int enable, long long *dst, size_t n) {
    if (!h) return fail(HIRL_E_INVALID, "null handle");
    const size_t cnt = (size_t)PROF_STEPS * ST_END * MAX_CTAS * 2;
    CU(cudaSetDevice(h->cfg.device));
    if (enable && !h->d_prof) {
        if (int rc = dev_alloc(h, &h->d_prof, cnt)) return rc;
    }
    h->prof_on = enable != 0;
    if (dst) {
        if (n < cnt || !h->d_prof) return fail(HIRL_E_INVALID, "profile buffer not enabled or destination too small");
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaMemcpy(dst, h->d_prof, cnt * 8, cudaMemcpyDeviceToHost));
    }
    return 0;
}
int hirl_debug_export(hirl_agent_t *h, int buffer_id, int pass, float *dst, size_t n, size_t *n_out) {
    if (!h) return fail(HIRL_E_INVALID, "null handle");
    const Scratch &s = h->ctxs[0].sc;
    const size_t B = h->cfg.batch, zs = h->ctxs[0].zsplit;
    const float *src = nullptr;
    size_t cnt = 0;
    switch (buffer_id) {
        case DB_X: src = s.X; cnt = B * ROW; break;
        case DB_XB: src = s.XB; cnt = B * BCW; break;
        case DB_H1: src = s.H1 + (size_t)pass * B * H1; cnt = B * H1; break;
        case DB_XH1: src = s.XH1 + (size_t)pass * B * H1; cnt = B * H1; break;
        case DB_RS1: src = s.RS1 + (size_t)pass * B; cnt = B; break;
        case DB_Z2: src = s.Z2 + (size_t)pass * zs * B * H2; cnt = zs * B * H2; break;
        case DB_DH1: src = s.DH1 + (size_t)pass * zs * B * H1; cnt = zs * B * H1; break;
        case DB_APRIME: src = s.APRIME; cnt = B * A; break;
        case DB_PI: src = s.PI; cnt = B * A; break;
        case DB_PIB: src = s.PIB; cnt = B * A; break;
        case DB_ABC: src = s.ABC; cnt = B * A; break;
        case DB_GPI: src = s.GPI; cnt = B * A; break;
        case DB_Y: src = s.Y; cnt = B; break;
        case DB_Q: src = s.Q; cnt = 2 * B; break;
        case DB_TQ: src = s.TQ; cnt = 2 * B; break;
        case DB_RLQ: src = s.RLQ; cnt = B; break;
        case DB_SOFTQ: src = s.SOFTQ; cnt = B; break;
        case DB_LP: src = s.LP; cnt = (size_t)MAX_CTAS * LP_WORDS; break;
        case DB_GW2P: src = s.GW2P; cnt = (size_t)2 * h->ctxs[0].dw_split * H2 * H1; break;
        case DB_XH2: src = s.XH2 + (size_t)pass * h->ctxs[0].Bpad * H2; cnt = (size_t)h->ctxs[0].Bpad * H2; break;
        case DB_RS2: src = s.RS2 + (size_t)pass * h->ctxs[0].Bpad; cnt = (size_t)h->ctxs[0].Bpad; break;
        case DB_DZ2S: src = s.DZ2S + (size_t)pass * cl::CL * h->ctxs[0].Bpad * cl::NS; cnt = (size_t)cl::CL * h->ctxs[0].Bpad * cl::NS; break;
        case DB_DZ1: src = s.DZ1 + (size_t)pass * B * H1; cnt = B * H1; break;
        case DB_DY1: src = s.DY1 + (size_t)pass * B * H1; cnt = B * H1; break;
        default: return fail(HIRL_E_INVALID, "bad debug buffer id");
    }
    if (n_out) *n_out = cnt;
    if (!dst) return 0;
    if (n < cnt) return fail(HIRL_E_INVALID, "destination too small");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    if (h->ctxs[0].use_cluster && buffer_id == DB_H1) {   // padded activation rows of the cluster kernel
        CU(cudaMemcpy2D(dst, H1 * 4, s.H1P + (size_t)pass * h->ctxs[0].Bpad * cl::WLD, cl::WLD * 4, H1 * 4, B, cudaMemcpyDeviceToHost));
        return 0;
    }
    CU(cudaMemcpy(dst, src, cnt * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- env -----------------------------------------------------------------------------------
static int env_grid(long long n, int per_block) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (n + per_block - 1) / per_block;
    return (int)std::max<long long>(1, std::min<long long>(tiles, (long long)sms * 32));
}
int hirl_env_step_dev(const float *raw16_dev, int64_t n, float *obs13_dev, float *reward_dev, int32_t *flags_dev,
                      unsigned long long *counters3_dev, void *stream) {
    if (n < 0 || (n > 0 && (!raw16_dev || !obs13_dev || !reward_dev || !flags_dev))) return fail(HIRL_E_INVALID, "bad argument");
    if (n == 0) return 0;
    if (((uintptr_t)raw16_dev | (uintptr_t)obs13_dev) & 15) return fail(HIRL_E_INVALID, "raw/obs pointers must be 16-byte aligned");
    env_step_kernel<<<env_grid(n, ENV_THREADS), ENV_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(raw16_dev), n, obs13_dev,
                                                                                         reward_dev, flags_dev, counters3_dev);
    CU(cudaGetLastError());
    return 0;
}
int hirl_env_step(int device, const float *raw16, int64_t n, float *obs13, float *reward, int32_t *flags, unsigned long long *counters3) {
    if (n < 0 || (n > 0 && (!raw16 || !obs13 || !reward || !flags))) return fail(HIRL_E_INVALID, "bad argument");
    if (n == 0) return 0;
    CU(cudaSetDevice(device));
    float *d_raw = nullptr, *d_obs = nullptr, *d_rew = nullptr;
    int32_t *d_fl = nullptr;
    unsigned long long *d_cnt = nullptr;
    CU(cudaMalloc(&d_raw, (size_t)n * 64));
    CU(cudaMalloc(&d_obs, (size_t)n * 52));
    CU(cudaMalloc(&d_rew, (size_t)n * 4));
    CU(cudaMalloc(&d_fl, (size_t)n * 4));
    if (counters3) { CU(cudaMalloc(&d_cnt, 24)); CU(cudaMemcpy(d_cnt, counters3, 24, cudaMemcpyHostToDevice)); }
    CU(cudaMemcpy(d_raw, raw16, (size_t)n * 64, cudaMemcpyHostToDevice));
    int rc = hirl_env_step_dev(d_raw, n, d_obs, d_rew, d_fl, d_cnt, nullptr);
    if (!rc) {
        CU(cudaMemcpy(obs13, d_obs, (size_t)n * 52, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(reward, d_rew, (size_t)n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(flags, d_fl, (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (counters3) CU(cudaMemcpy(counters3, d_cnt, 24, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_raw); cudaFree(d_obs); cudaFree(d_rew); cudaFree(d_fl); cudaFree(d_cnt);
    return rc;
}
int hirl_env_offline_dev(const float *s13_dev, const float *a4_dev, const float *s2_13_dev, int64_t n, float *reward_dev, int8_t *success_dev,
                         uint8_t *done_dev, void *stream) {
    if (n < 0 || (n > 0 && (!s13_dev || !a4_dev || !s2_13_dev || !reward_dev || !success_dev || !done_dev))) return fail(HIRL_E_INVALID, "bad argument");
    if (n == 0) return 0;
    env_offline_kernel<float><<<env_grid(n, ENV_THREADS), ENV_THREADS, 0, (cudaStream_t)stream>>>(s13_dev, a4_dev, s2_13_dev, n, reward_dev,
                                                                                            reinterpret_cast<signed char *>(success_dev), done_dev);
    CU(cudaGetLastError());
    return 0;
}
}  // extern "C"
// host-buffer form of the offline pass for fp32 rows (T = float) and for the float64 expert table (T = double)
template <typename T>
static int env_offline_host(int device, const T *s13, const T *a4, const T *s2_13, int64_t n, float *reward, int8_t *success, uint8_t *done) {
    if (n < 0 || (n > 0 && (!s13 || !a4 || !s2_13 || !reward || !success || !done))) return fail(HIRL_E_INVALID, "bad argument");
    if (n == 0) return 0;
    CU(cudaSetDevice(device));
    T *d_s = nullptr, *d_a = nullptr, *d_s2 = nullptr;
    float *d_r = nullptr;
    int8_t *d_su = nullptr;
    uint8_t *d_d = nullptr;
    const size_t w = sizeof(T);
    CU(cudaMalloc(&d_s, (size_t)n * 13 * w)); CU(cudaMalloc(&d_a, (size_t)n * 4 * w)); CU(cudaMalloc(&d_s2, (size_t)n * 13 * w));
    CU(cudaMalloc(&d_r, (size_t)n * 4)); CU(cudaMalloc(&d_su, (size_t)n)); CU(cudaMalloc(&d_d, (size_t)n));
    CU(cudaMemcpy(d_s, s13, (size_t)n * 13 * w, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_a, a4, (size_t)n * 4 * w, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_s2, s2_13, (size_t)n * 13 * w, cudaMemcpyHostToDevice));
    env_offline_kernel<T><<<env_grid(n, ENV_THREADS), ENV_THREADS>>>(d_s, d_a, d_s2, n, d_r, reinterpret_cast<signed char *>(d_su), d_d);
    CU(cudaGetLastError());
    CU(cudaMemcpy(reward, d_r, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(success, d_su, (size_t)n, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(done, d_d, (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(d_s); cudaFree(d_a); cudaFree(d_s2); cudaFree(d_r); cudaFree(d_su); cudaFree(d_d);
    return 0;
}
// train_all.py:287-306.  The reward/flag pass over all consecutive pairs runs on the GPU; the
// data-dependent "skip the row after a terminal next-state" walk is a sequential scan over n
// bytes on the host (it has a loop-carried dependence and costs microseconds).  Stored rows are the fp32 casts
// the reference makes at sample time (HIRL.py:229-233), whatever the input precision.
template <typename T>
static int expert_scan_host(int device, const T *states13, const T *actions4, int n, float *out_rows32, int8_t *out_success, int *n_rows_out) {
    if (!states13 || !actions4 || !out_rows32 || !out_success || !n_rows_out || n < 0) return fail(HIRL_E_INVALID, "bad argument");
    *n_rows_out = 0;
    if (n < 2) return 0;
    const int m = n - 1;
    std::vector<float> r(m);
    std::vector<int8_t> su(m);
    std::vector<uint8_t> d(m);
    if (int rc = env_offline_host<T>(device, states13, actions4, states13 + S, m, r.data(), su.data(), d.data())) return rc;
    int kept = 0;
    for (int i = 0; i < m;) {
        float *row = out_rows32 + (size_t)kept * ROW;
        for (int k = 0; k < S; ++k) row[k] = (float)states13[(size_t)i * S + k];
        for (int k = 0; k < A; ++k) row[S + k] = (float)actions4[(size_t)i * A + k];
        for (int k = 0; k < S; ++k) row[SA + k] = (float)states13[(size_t)(i + 1) * S + k];
        row[30] = r[i];
        row[31] = d[i] ? 1.f : 0.f;
        out_success[kept] = su[i];
        ++kept;
        i += d[i] ? 2 : 1;
    }
    *n_rows_out = kept;
    return 0;
}
extern "C" {
int hirl_env_offline(int device, const float *s13, const float *a4, const float *s2_13, int64_t n, float *reward, int8_t *success, uint8_t *done) {
    return env_offline_host<float>(device, s13, a4, s2_13, n, reward, success, done);
}
int hirl_env_offline_f64(int device, const double *s13, const double *a4, const double *s2_13, int64_t n, float *reward, int8_t *success, uint8_t *done) {
    return env_offline_host<double>(device, s13, a4, s2_13, n, reward, success, done);
}
int hirl_expert_scan(int device, const float *states13, const float *actions4, int n, float *out_rows32, int8_t *out_success, int *n_rows_out) {
    return expert_scan_host<float>(device, states13, actions4, n, out_rows32, out_success, n_rows_out);
}
int hirl_expert_scan_f64(int device, const double *states13, const double *actions4, int n, float *out_rows32, int8_t *out_success, int *n_rows_out) {
    return expert_scan_host<double>(device, states13, actions4, n, out_rows32, out_success, n_rows_out);
}

int hirl_measure_fp32_peak(int device, double *tflops_out) {
    if (!tflops_out) return fail(HIRL_E_INVALID, "null argument");
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    float *d = nullptr;
    CU(cudaMalloc(&d, 4));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int iters = 1 << 14, blocks = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CU(cudaEventRecord(e0));
        fma_peak_kernel<<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 16 * iters * 256.0 * blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops_out = best;
    return 0;
}

}  // extern "C"
