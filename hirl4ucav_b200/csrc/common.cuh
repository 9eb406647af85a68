// common.cuh -- layouts and the kernel context shared by the update kernels and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hirl_b200.h"

namespace hirl {

constexpr int S = HIRL_STATE_DIM;    // 13
constexpr int A = HIRL_ACTION_DIM;   // 4
constexpr int SA = S + A;            // 17
constexpr int H1 = HIRL_H1;          // 256
constexpr int H2 = HIRL_H2;          // 512
constexpr int ROW = HIRL_ROW_WORDS;  // 32
constexpr int BCW = HIRL_BC_WORDS;   // 20
constexpr float LN_EPS = 1e-5f;      // nn.LayerNorm default (reference HIRL.py:28)

// Internal tower layout (floats).  Every segment starts on a 32-float (128 B) boundary so all
// vector loads are aligned; W1 is kept TRANSPOSED ([k][n]) because the first layer is evaluated
// one row per warp with lanes striding over n.  import/export convert from/to state_dict order.
template <int K0, int NOUT>
struct TowerLayout {
    static constexpr int k0 = K0, nout = NOUT;
    static constexpr int W1T = 0;                  // [K0][H1]
    static constexpr int B1 = W1T + K0 * H1;       // [H1]
    static constexpr int G1 = B1 + H1;             // layernorm1.weight
    static constexpr int BE1 = G1 + H1;            // layernorm1.bias
    static constexpr int W2 = BE1 + H1;            // [H2][H1] row-major (out,in)
    static constexpr int B2 = W2 + H2 * H1;
    static constexpr int G2 = B2 + H2;
    static constexpr int BE2 = G2 + H2;
    static constexpr int W3 = BE2 + H2;            // [NOUT][H2]
    static constexpr int B3 = W3 + NOUT * H2;      // [NOUT] padded to 32
    static constexpr int SIZE = B3 + 32;
    static constexpr int SMALL = SIZE - H2 * H1;   // everything but W2 (per-CTA partial grads)
    // index into the "small" partial-gradient layout for an arena offset outside W2
    __host__ __device__ static constexpr int small_index(int off) { return off < W2 ? off : off - H2 * H1; }
};
using ActorL = TowerLayout<S, A>;    // 138784 floats
using CriticL = TowerLayout<SA, 1>;  // 138272 floats per tower, two towers per critic
static_assert(ActorL::SIZE % 32 == 0 && CriticL::SIZE % 32 == 0, "tower arenas must stay 128B aligned");

constexpr int ACTOR_SIZE = ActorL::SIZE;
constexpr int CRITIC_SIZE = 2 * CriticL::SIZE;

// forward passes of one learn() call (scratch buffers are indexed by these)
enum Pass { P_TA = 0, P_C1, P_C2, P_TC1, P_TC2, P_A, P_AB, P_BC, P_QP, P_QS, N_PASS };
// passes that are back-propagated keep xhat1/rstd1 and receive dH1: index into those arrays
__host__ __device__ constexpr int bwd_slot(int pass) {
    return pass == P_C1 ? 0 : pass == P_C2 ? 1 : pass == P_A ? 2 : pass == P_AB ? 3 : pass == P_QP ? 4 : -1;
}
constexpr int N_BWD = 5;
constexpr int MAX_SPLIT = 4;   // split-K partial buffers per GEMM output
// split-K factors are chosen per GEMM group on the host (makespan heuristic); producers and
// consumers of a buffer read the same entry.
enum SplitIdx { NS_G1 = 0, NS_G1_ACTOR, NS_G2, NS_G3DX, NS_G5, NS_G6, NS_G7DX, NS_COUNT };
constexpr int MAX_CTAS = 160;  // >= SM count (148 on B200)
constexpr int LP_WORDS = 16;   // per-CTA loss/count partial words

// stages of one learn() call, in execution order (hirl_debug_stop_after uses these numbers)
enum Stage {
    ST_R0 = 1,  // gather + first layers (TA, C1, C2 [+ A, AB, BC])
    ST_G1,      // second-layer GEMMs of those passes
    ST_R1,      // TA head -> a' ; first layers of TC1, TC2 ; [actor heads -> pi, pi_bc, a_bc]
    ST_G2,      // second-layer GEMMs of TC1, TC2
    ST_R2,      // TD target, critic heads, loss, backward through head + LayerNorm2
    ST_G3,      // dW2 and dH1 of both critic towers
    ST_R3,      // backward through LayerNorm1 / first layer of the critic
    ST_G4,      // critic Adam
    ST_R4,      // first layers of Q1(s,pi), Q1(s,a_bc) with the UPDATED critic
    ST_G5,      // their second-layer GEMMs
    ST_R5,      // Q heads, soft count, backward of Q1(s,pi) to dz2
    ST_G6,      // dH1 of Q1(s,pi)
    ST_R6,      // d pi ; actor heads backward (both branches) to dz2
    ST_G7,      // actor dW2, dH1 (both branches)
    ST_R7,      // backward through LayerNorm1 / first layer of the actor
    ST_G8,      // actor Adam, Polyak, scalars
    ST_END
};

// debug buffer ids (hirl_debug_export)
enum DebugBuf { DB_X = 0, DB_XB, DB_H1, DB_XH1, DB_RS1, DB_Z2, DB_DH1, DB_APRIME, DB_PI, DB_PIB, DB_ABC,
                DB_Y, DB_Q, DB_TQ, DB_RLQ, DB_SOFTQ, DB_GPI, DB_LP, DB_GW2P, DB_XH2, DB_RS2, DB_DZ2S, DB_DZ1, DB_DY1 };

struct Scratch {
    float *X;        // [B][32] gathered replay rows
    float *XB;       // [B][20] gathered expert (s,a) rows
    float *H1;       // [N_PASS][B][H1]
    float *XH1;      // [N_BWD][B][H1]
    float *RS1;      // [N_BWD][B]
    float *Z2;       // [N_PASS][nsplit][B][H2]   second-layer pre-activations, later dz2 (split 0)
    float *DH1;      // [N_BWD][nsplit][B][H1]
    float *APRIME;   // [B][4]
    float *PI, *PIB, *ABC, *GPI;  // [B][4]
    float *Y;        // [B]
    float *Q;        // [2][B]
    float *TQ;       // [2][B]
    float *RLQ, *SOFTQ;           // [B]
    float *PART_C;   // [MAX_CTAS][2][CriticL::SMALL]
    float *PART_A;   // [MAX_CTAS][ActorL::SMALL]
    float *LP;       // [MAX_CTAS][LP_WORDS]
    float *GW2P;     // [2 * dw_split][H2][H1] split-K / per-branch partial dW2 buffers, summed in fixed order by Adam
    float *H1T;      // [4][H1][BT]  transposed h1 of C1, C2, A, AB   (tensor-core dW operands, K-major)
    float *DZ2T;     // [4][H2][BT]  transposed dz2 of the same passes
    // cluster-resident kernel (csrc/cluster_update.cuh)
    float *H1P;      // [N_PASS][Bpad][H1 + 4]  first-layer activations in padded rows: a row block is one contiguous bulk copy
    float *DZ2S;     // [4][8 slices][Bpad][64] dz2 of C1, C2, A, AB, slice-major: a dW2 tile's operand is one contiguous block
    float *DZ1, *DY1;  // [4][B][H1]  gradients at the first layer's pre-activation / LayerNorm1 output
    float *XH2;      // [2][Bpad][H2] xhat2 of passes A, AB (kept from phase A for the actor backward)
    float *RS2;      // [2][Bpad]
    float *PART2C;   // [clusters][PC_WORDS] per-cluster column sums of the critic's small second-layer parameters
    float *PART2A;   // [clusters][PA_WORDS] same for the actor, per branch
};

// Tensor-core cluster kernel (csrc/tcc_update.cuh): operand images of every W2 and its scratch
struct TccPtrs {
    // forward images: per tower [8 feature slices][8 k-blocks][hi 2048 | lo 2048 floats], SWIZZLE_128B tiles of [64][32]
    float *img_actor, *img_tactor, *img_bc, *img_critic, *img_tcritic;
    // input-gradient images (W2 read along its output features): per tower [8 slices][2 m-blocks][2 k-chunks][hi 4096 | lo 4096]
    float *imgt_actor, *imgt_critic;
    float *DH1P;     // [units][8 CTAs][64 rows][H1]  K-split partials of dH1
    float *GW2PC;    // [2 towers][4 row blocks][H2][H1] per-row-block dW2 of the critic
    float *GW2PA;    // [2 branches * 4 row blocks][H2][H1] of the actor
    float *PARTC;    // [4 row blocks * 8 CTAs][2 towers][CriticL::SMALL] small-parameter partials
    float *PARTA;    // [2 branches * 4 row blocks * 8 CTAs][ActorL::SMALL]
    float *TQX;      // [4][2][64] target Q of a tower, exchanged between the two tower clusters of a row block
    float *RLQX;     // [4][64] Q1(s, pi) of a row block, published by role A for role S' count
    float *PIX;      // [4][64][4] pi = actor(s), evaluated by phase A's actor units
    float *RSA;      // [4][64] rstd of LayerNorm2 of that pass
    float *OWNP;     // [4][64][H1 + 32] xhat1 (+ rstd1) of that pass
    float *APX;      // [4][64][4] a' of a row block, published by its target-actor unit
    float *XHP;      // [8 units][64][H2] normalised second-layer rows of the critic forward pass, parked while the target tower runs
    float *XH2S;     // [4][64][H2] normalised second-layer rows of the actor pass, parked while the Q pass runs
    float *H1I;      // [clusters][32768] first-layer activation tile (hi | lo) of the pass a cluster is executing: written row-wise by its 8 CTAs, multicast back
    float *H1TI;     // [units][32768] transposed first-layer activation tiles (hi | lo) of back-propagated passes
    float *RSUM;     // [units][8] per-unit row sums (scratch)
    float *LPT;      // [16] per-unit loss partials: critic [r][t] at 0..7, cloning (sum sq, fire sq) at 8 + 2 r
    float *CNT;      // [4][2] soft count and sum of rl_Q per row block
    unsigned *pflags;   // [16][32] point-to-point flags (generations keep counting across launches)
    unsigned *gen_ctr;  // calls executed so far by this agent: base of the flag generations of a launch
};

// Everything one replica's update needs.  Passed to the kernel by value (per-replica array in
// global memory for ensembles).
struct AgentCtx {
    // parameters (internal layout), gradients and Adam moments
    float *actor, *tactor, *critic, *tcritic, *bc;
    float *g_actor, *g_critic;
    float *m_actor, *v_actor, *m_critic, *v_critic;
    float *w2t_actor, *w2t_critic;   // transposed copies of W2 ([H1][H2], per tower): K-major operand of the tensor-core dX GEMMs
    // padded images of every W2 ([H2][H1 + 4] per tower): the cluster kernel streams a CTA's 64-row slice with one bulk copy
    float *w2p_actor, *w2p_tactor, *w2p_bc, *w2p_critic, *w2p_tcritic;
    // replay tables
    const float *online;   // [online_capacity][32]
    const float *expert;   // [expert_capacity][32]
    const float *bctab;    // [bc_rows][20]
    Scratch sc;
    // per-launch step streams
    const hirl_step_t *steps;
    const int32_t *idx_online, *idx_expert, *idx_bc;  // [k][B]
    float *out_scalars;    // [k][8]
    float *carry;          // [8] scalars carried between calls (stale-value semantics, HIRL.py:334)
    int B;
    int Bg;                // batch the loss means run over: B, or the global batch in the data-parallel mode (this agent holds B of its rows)
    float *dp_count;       // data-parallel mode, NCCL variant: soft-count word, written by the launch that ends with stage G6, all-reduced, read by stage R6 (else null)
    // data-parallel mode, FUSED variant (in-kernel all-reduce over NVLink peer memory, update.cuh: dp_*): every rank's exchange
    // block (layout DPX_*), own block included, mapped into this process with CUDA IPC; 0 ranks = off
    int dp_world, dp_rank;
    unsigned *dp_gen_ctr;  // calls executed so far by this agent (device word; the flag generations of a launch start here)
    char *dp_peer[8];
    int zsplit;            // allocated split-K slots per Z2 / DH1 buffer (max of ns[])
    int ns[NS_COUNT];      // split-K factor per GEMM group
    int activation;
    int use_tc;            // 1: 256<->512 contractions on tcgen05 (3xTF32), 0: FP32-FMA tiles
    int tc_g1;             // 1: only the forward GEMMs of stage G1 on tcgen05 (mid-size batches, see hirl_agent_create)
    int dw_split;          // split-K factor of the dW2 GEMMs (reduction over the batch), >= 1
    int BT;                // batch rounded up to 32: leading dimension of the transposed buffers
    int use_tcc;           // 1: tensor-core cluster kernel (csrc/tcc_update.cuh)
    TccPtrs tcc;
    int use_cluster;       // 1: cluster-resident kernel (csrc/cluster_update.cuh); rows per block = 2 * cluster_ri
    int cluster_ri, Bpad;  // Bpad = row blocks * rows per block
    float gamma, tau, noise_clamp, loss_lambda;
};

// Exchange block of the fused data-parallel mode (floats from the block base).  Gradients and scalars are double-buffered by
// call parity: a rank overwrites a buffer only two all-reduces after its peers read it, and nobody can be two all-reduces
// ahead of anybody (every all-reduce waits for every rank).  Flags are generations (call numbers), never reset.
constexpr size_t DPX_GC = 0;                                     // [2][CRITIC_SIZE] this rank's critic gradient of its shard
constexpr size_t DPX_GA = DPX_GC + 2 * (size_t)CRITIC_SIZE;      // [2][ACTOR_SIZE]
constexpr size_t DPX_SC = DPX_GA + 2 * (size_t)ACTOR_SIZE;       // [2][32]: 0 critic loss, 1..4 actor scalars (carry[1..4]), 5 soft count
constexpr size_t DPX_FLAGS = DPX_SC + 64;                        // unsigned [3][32]: critic gradient / count / actor gradient ready
constexpr size_t DPX_FLOATS = DPX_FLAGS + 3 * 32;
enum { DPF_CRITIC = 0, DPF_COUNT = 1, DPF_ACTOR = 2 };

struct LaunchCtx {
    AgentCtx agent;        // single-agent launch: the context travels as a kernel parameter
    unsigned *bar_flags;   // [MAX_CTAS][32] barrier flags: generations keep counting up across launches (never reset)
    int k_steps;
    int stop_after;        // debug: leave after this stage of step 0 (0 = never)
    int step0;             // first step record of this launch
    int first_stage, last_stage, adam_phase;   // stage range of a data-parallel launch (1, ST_END, 0: whole calls)
    long long *prof;       // debug: [PROF_STEPS][ST_END][MAX_CTAS][2] SM-clock stamps (work done, barrier released)
};
constexpr int PROF_STEPS = 12;

// Ensemble launch: R independent replicas processed by n_teams teams of team_size CTAs; replica r is
// handled by team r % n_teams (its AgentCtx.sc points at that team's scratch).
struct EnsembleLaunch {
    const AgentCtx *agents;  // [replicas] in global memory
    unsigned *bar_flags;     // [n_teams][team_size][32]
    int replicas, n_teams, team_size, k_steps;
};

}  // namespace hirl
