// Matrix-free covariance-kernel operator  W[I, :] = C[I, :] * X   (SURVEY.md §8 a8).
//
// C[i,j] = sigma2 * k(r2(u_i, u_j)) + nugget * (i == j) is never materialised.  Per
// 32-point k-tile, the four warps that share a 16-row group generate that group's
// 16 x 32 block of kernel values ONCE (4 entries per thread, FP64 DFMA/exp) into a
// double-buffered shared-memory tile laid out for conflict-free DMMA A-fragment loads,
// synchronise through a per-row-group mbarrier (arrive after publishing, wait right before
// the first read, so stragglers are covered by the next tile's generation work), and then each warp multiplies it against its
// quarter of the X tile with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64 -> DMMA.8x8x4),
// accumulating a 16 x (8*NB/4) strip in registers.  X tiles (TALL layout, a k-tile is one
// contiguous chunk) and the scaled coordinates arrive by 1-D bulk async copies (UBLKCP)
// completing on mbarriers; thread 0 issues them `lookahead` tiles ahead.
//
// Why this shape (profiles/r01): a single warp can issue a DMMA only every ~32 cycles
// while the pipe retires one per 16, and DFMA shares that pipe -- so the SM needs >= 3-4
// warps per scheduler (16 warps/SM, <= 128 registers each) and every kernel value must be
// generated exactly once per CTA.  The 8-warp / A-in-registers first version stalled 43 %
// of samples in `wait` with the DMMA pipe 71 % busy.
//
// CTA = 16 warps = 4 row groups (16 rows) x 4 column groups: 64 rows x 8*NB columns.
// Persistent grid: one CTA per SM (x occupancy), static schedule: full rounds of 64-row
// tiles plus one tail round at 16-row-group granularity (see the schedule comment below).
#include "common.cuh"
#include "algos.h"
#include "ptx.cuh"
#include "nb_list.h"

namespace gsi {

constexpr int KC_BM = 64;          // rows per CTA tile
constexpr int KC_BK = 32;          // j-points (GEMM K) per pipeline stage
constexpr int KC_AP = KC_BK + 4;   // pitch of the generated A tile (4 mod 16 doubles)
constexpr int KC_WARPS = 16;
constexpr int KC_THREADS = KC_WARPS * 32;
constexpr int KC_CG = 4;           // column groups (warps sharing one row group)
constexpr int KC_KIND_TABLE = 3;   // internal kind: stationary kernel on a structured grid, values from a lattice table
constexpr long long KC_SPIN_LIMIT = 400000;   // cycles (~0.2 ms) a producer waits on the sweep window before giving it up

struct KcovParams {
    const double* X;       // TALL, all n rows (zero padded), pitch ld
    double* W;             // TALL, local rows, pitch ldw
    const double* u;       // scaled coordinates [3][n_pad]              (arithmetic generation)
    const int* lat;        // lattice indices [3][n_pad]                 (structured grid: table lookup)
    const double* table;   // k(r2) for every lattice offset: table[dx + nx*(dy + ny*dz)]
    int nx, ny;
    int sweep_groups, sweep_div, l2_hint;   // k-sweep de-synchronisation (power-of-two groups, spread = groups/div of X)
    int pad0;
    unsigned int* sync_cnt;                 // sweep window: arrivals per epoch (zeroed before the launch)
    int win_epochs, epoch_shift;            // a CTA runs at most win_epochs epochs of 2^epoch_shift k-tiles ahead of the slowest
    int64_t n;             // columns of C (= rows of X)
    int64_t n_pad;
    int64_t row0;          // first global row of this rank's block
    int64_t mloc;          // local rows
    int64_t ld, ldw;
    double sigma2, nugget, beta;
    int stages;
    // stream-K tail (see the schedule comment in the kernel): the 64-row tiles left over after the full rounds
    // are split along k into equal shares of tail_u k-tiles per CTA; partial sums go to `partial`
    int64_t tail_u;        // k-tiles of tail work per CTA (0: no tail)
    double* partial;       // [2 * grid][KC_BM][ldw] raw accumulators of the tail segments (slot 2b + s)
};

// 2^(j/64), j = 0..63, correctly rounded (generated with 200-bit arithmetic)
__constant__ double kExp2Table[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

// exp(-y) for y >= 0 with a 64-entry table and a degree-5 polynomial: 10 FP64 pipe
// instructions instead of the ~17 of the library exp() -- the kernel generation shares the
// FP64 pipe with the DMMAs (a register-resident DMMA loop drops from 37.1 to 31.6 TF/s when
// one library-exp Gaussian value is generated per 28 DMMAs, tools/mma_lds_peak.cu), so
// every DFMA removed here is tensor throughput.  Max error 2 ulp (library exp: 1 ulp);
// results below 2^-1000 are flushed to zero.  `tab` is the table staged in shared memory.
__device__ __forceinline__ double fast_exp_neg(double y, const double* __restrict__ tab) {
    const double L = 92.33248261689366;                    // 64 / ln 2
    const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52: round-to-nearest integer trick
    const double LN2_64_HI = 0x1.62e42fef80000p-7;         // ln2/64, 34 significant bits (k * HI exact)
    const double LN2_64_LO = 0x1.1cf79abc9e3b4p-42;
    const double t = fma(y, -L, MAGIC);
    const int k = __double2loint(t);                       // k = round(-y * 64 / ln2) <= 0
    const double kf = t - MAGIC;
    double r = fma(kf, -LN2_64_HI, -y);
    r = fma(kf, -LN2_64_LO, r);                            // |r| <= ln2/128
    double p = 1.0 / 120.0;
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = tab[k & 63] * p;                      // in [1/2 .. 2)
    const int e = k >> 6;                                  // floor(k / 64) <= 0
    const int hi = __double2hiint(v) + (e << 20);
    const double res = __hiloint2double(hi, __double2loint(v));
    // y >= 1024 (hi word >= 0x40900000; also inf/NaN): k would wrap for y >~ 2.3e7 -- flush by the
    // integer compare on y itself (no FP64-pipe instruction)
    return (e < -1000 || __double2hiint(y) >= 0x40900000) ? 0.0 : res;
}

// sqrt(x) for x > 0 (callers add 1e-300 so the diagonal r2 = 0 stays finite): hardware
// rsqrt seed (MUFU, 2^-22) + two Goldschmidt steps = 6 FP64 pipe instructions, <= 2 ulp.
__device__ __forceinline__ double fast_sqrt_pos(double x) {
    double rs;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(x));
    double h = __hiloint2double(__double2hiint(rs) - 0x00100000, __double2loint(rs));   // rs / 2
    double g = x * rs;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    return g;
}

// kind-specific value from r2.  For the Gaussian kernel the coordinates handed to the
// kernel are pre-scaled by 1/sqrt(2) so that r2 already carries the factor 1/2.
template <int KIND>
__device__ __forceinline__ double kern_eval(double r2, double beta, const double* __restrict__ tab) {
    if (KIND == GSI_KERNEL_EXPONENTIAL) return fast_exp_neg(fast_sqrt_pos(r2), tab);
    if (KIND == GSI_KERNEL_GAUSSIAN) return fast_exp_neg(r2, tab);
    if (KIND == KC_KIND_TABLE) return 0.0;
    return exp(-beta * log1p(r2));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// (Two probes of the first round -- the X tile fetched as four paced bulk copies, and an L2 bulk
// prefetch a few k-tiles ahead of every CTA's sweep -- were run on hardware in round 2
// (profiles/r02/l2_question.md): both bit-identical, the paced fetch 19 % slower, the prefetch
// within 0.1 % of the default although it turns every shared-memory fill into an L2 hit.  So an
// L2-served X stream is not slower as such; both variants were removed.  A third probe, round 2: the tile fetched
// as 4 / 8 / 32 bulk copies in an order rotated by the CTA index -- the idea being that CTAs sweeping X in step
// (sweep window) hammer the same L2 lines in the same order -- made every schedule slower (windowed front 563 ->
// 623 / 652 / 836 ms, default 516 -> 559 ms: the per-copy cost outweighs any spreading; gpurun_out r03_chunks.log);
// removed as well.)
template <int NB, int KIND, int DIM>
__global__ void __launch_bounds__(KC_THREADS, 1) kcov_gemm_kernel(const __grid_constant__ KcovParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ld = NB * 8 + 4;
    constexpr int NBW = (NB + KC_CG - 1) / KC_CG;                // n-blocks per warp
    constexpr int stage_doubles = KC_BK * ld + 3 * KC_BK;        // X tile + coordinate tile
    constexpr int a_doubles = KC_BM * KC_AP;
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* a_tiles = smem + (size_t)p.stages * stage_doubles;   // [2][KC_BM][KC_AP]
    uint64_t* full = reinterpret_cast<uint64_t*>(a_tiles + 2 * a_doubles);
    uint64_t* empty = full + p.stages;
    uint64_t* abar = empty + p.stages;                            // one per row group, 4 warp arrivals
    double* etab = reinterpret_cast<double*>(abar + 4);           // 2^(j/64) table for fast_exp_neg

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int nstages = p.stages;
    const int rg = warp >> 2;                 // row group: rows rg*16 .. rg*16+15 of the tile
    const int cg = (warp + rg) & 3;           // column group, rotated per row group so that each SM
                                              // sub-partition (warp % 4) hosts all four column groups
                                              // (the last group may own fewer n-blocks)
    const int nb0 = cg * NBW;                 // first n-block of this warp

    if (tid < 64) etab[tid] = kExp2Table[tid];
    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], KC_WARPS);
        }
        for (int r = 0; r < 4; ++r) mbar_init(&abar[r], KC_CG);
        mbar_fence_init();
    }
    __syncthreads();

    // Work schedule.  Full rounds: CTA b takes the 64-row tile (round*grid + b), i.e. four consecutive
    // 16-row groups, and sweeps all nkt k-tiles (cyclically, from its own start tile kt0).  The tiles left
    // over after the full rounds (< grid of them) form the STREAM-K TAIL: their tail_tiles * nkt k-tile
    // units are cut into equal contiguous shares of p.tail_u units, one share per CTA, so a share is a k
    // range of one tile or the end of one tile plus the start of the next -- at most two SEGMENTS.  A tail
    // segment stores its raw accumulators to p.partial (slot 2b + s); kcov_tail_fixup_kernel adds the
    // pieces of a tile in ascending CTA order (deterministic) and applies sigma2 / nugget.  Every SM thus
    // ends within one k-tile of every other whatever the row count (round 1 spread the leftover 16-row
    // groups over the CTAs with full k sweeps: 4.5 % of a C3 product at 8 GPUs, where a rank's 1568 row
    // groups are 2.65 rounds).
    const int64_t total_rg = (p.mloc + 15) / 16;
    const int64_t per_round = (int64_t)gridDim.x * 4;
    const int64_t full_rounds = total_rg / per_round;
    const int64_t tail_rg0 = full_rounds * per_round;                         // first row group of the tail
    const int64_t tail_tiles = (total_rg - tail_rg0 + 3) / 4;
    const int64_t nkt = (p.n + KC_BK - 1) / KC_BK;
    // De-synchronised (cyclic) k sweeps (gsi_ctx_set_option "kcov.sweep_*" / GSI_SWEEP): CTA b starts
    // its sweep (b mod groups) * separation tiles into X (default 64 groups over a quarter of X).
    // Measured on C3 (profiles/r01): +2.5 % over lock-step starts.  In this mode the X stream
    // (366 MB > L2) is served mostly from HBM -- ncu: 610-716 GB read per launch (L2 hit rate
    // 26-33 %, 17 % of HBM bandwidth) -- with the DMMA pipe 91-93 % busy, 3.95 us per k-tile.
    const int64_t kt_sep = p.sweep_div < 0 ? -(int64_t)p.sweep_div
                           : p.sweep_div == 0 ? 0
                           : (nkt >= p.sweep_div ? nkt / p.sweep_div : (nkt >= 16 ? 1 : 0));
    const int64_t kt0 = ((int64_t)(blockIdx.x & (p.sweep_groups - 1)) * kt_sep) % nkt;
    // Sweep window (p.win_epochs > 0, default off): the persistent CTAs all stream the same X, but
    // left alone they drift apart by more than an L2's worth of k-tiles.  The schedule is static --
    // the launch ends with its slowest CTA anyway -- so the producer can hold a CTA that is more than
    // win_epochs epochs ahead of the slowest one: every 2^epoch_shift k-tiles it counts itself into
    // the epoch's arrival counter and waits until the epoch win_epochs back has been reached by every
    // CTA that will ever reach it (all CTAs during the full rounds, the tail CTAs afterwards).  The
    // wait is bounded (KC_SPIN_LIMIT): a CTA that times out -- co-tenancy, a debugger -- drops the
    // window for the rest of the launch, so the window can cost time but never progress.
    // MEASURED (profiles/r01/sweep_schedules_c3_summary.csv, 22 schedules): one coherent front cuts
    // DRAM traffic 611 GB -> 8.0 GB per launch (= X once per round, L2 hit rate 98.6 %) and products
    // stay bit-identical, but EVERY L2-served variant -- any window, lock-step or de-phased CTAs,
    // 2..32 fronts, evict_last -- runs at 589-612 ms (4.5 us per k-tile) against 528-532 ms streamed
    // from HBM, and so does a problem whose X fits L2 outright; the pacing itself is free (same
    // window over the quarter-of-X spread: 530 ms).  The HBM-streamed schedule therefore stays the
    // default; the window is the knob for deployments that must spare HBM bandwidth.  Why the
    // L2-served stream costs the DMMA pipe ~13 % is the first ncu question of the next round.
    const int64_t full_it = full_rounds * nkt;
    // my share of the tail: units [u0, u1) -> segment A = (tile tA, k-tiles [kA, kA + lenA)), segment B = (tile tA + 1, [0, lenB))
    const int64_t tail_units = tail_tiles * nkt;
    int64_t u0 = (int64_t)blockIdx.x * p.tail_u, u1 = u0 + p.tail_u;
    if (u0 > tail_units) u0 = tail_units;
    if (u1 > tail_units) u1 = tail_units;
    const int64_t tA = u0 / nkt, kA = u0 - tA * nkt;
    const int64_t lenA = (u1 - u0 < nkt - kA) ? u1 - u0 : nkt - kA;
    const int64_t lenB = (u1 - u0) - lenA;
    const int nseg_tail = (lenA > 0 ? 1 : 0) + (lenB > 0 ? 1 : 0);
    const int64_t nseg = full_rounds + nseg_tail;
    // segment j: first row group, active row groups, first k-tile, k-tiles, partial slot (-1: direct output)
    struct Seg { int64_t base_rg; int nact; int64_t kbeg; int64_t klen; int slot; };
    auto seg_info = [&](int64_t j) -> Seg {
        Seg g;
        if (j < full_rounds) { g.base_rg = (j * gridDim.x + blockIdx.x) * 4; g.nact = 4; g.kbeg = 0; g.klen = nkt; g.slot = -1; return g; }
        const int sidx = (int)(j - full_rounds);
        const int64_t tile = tA + sidx;
        g.base_rg = tail_rg0 + tile * 4;
        const int64_t left = total_rg - g.base_rg;
        g.nact = (int)(left < 4 ? left : 4);
        g.kbeg = sidx == 0 ? kA : 0;
        g.klen = sidx == 0 ? lenA : lenB;
        g.slot = 2 * (int)blockIdx.x + sidx;
        return g;
    };
    // iteration `it` of this CTA -> its segment and global k-tile
    auto locate = [&](int64_t it_, int64_t& j, int64_t& kt) {
        if (it_ < full_it) {
            j = it_ / nkt;
            kt = it_ - j * nkt + kt0;
            if (kt >= nkt) kt -= nkt;
        } else {
            const int64_t r = it_ - full_it;
            if (r < lenA) { j = full_rounds; kt = kA + r; }
            else { j = full_rounds + 1; kt = r - lenA; }
        }
    };
    // CTAs that execute iteration x of the tail (x >= 0): those whose share is longer than x
    auto tail_ctas_beyond = [&](int64_t x) -> unsigned {
        if (p.tail_u <= 0 || x >= p.tail_u) return 0u;
        int64_t c = (tail_units - x + p.tail_u - 1) / p.tail_u;
        if (c > (int64_t)gridDim.x) c = gridDim.x;
        return c < 0 ? 0u : (unsigned)c;
    };
    bool window_on = p.win_epochs > 0;            // thread 0 only
    const uint64_t xpolicy = l2_policy_evict_last();
    constexpr uint32_t stage_bytes = (uint32_t)(KC_BK * ld * sizeof(double) +
                                                DIM * KC_BK * (KIND == KC_KIND_TABLE ? sizeof(int) : sizeof(double)));

    // ---------------- producer (thread 0): streams X / coordinate tiles ----------------------
    const int64_t total_it = full_it + (u1 - u0);
    const int lookahead = nstages > 2 ? nstages - 2 : 1;
    // Producer state, advanced incrementally (no 64-bit divisions on thread 0's path: its warp is an MMA
    // warp too, and what the producer spends per k-tile is added to the whole CTA's iteration time):
    //   p_it     next iteration to produce      p_s / p_ph   its pipeline stage / mbarrier phase
    //   p_kt     its global k-tile              p_left       k-tiles left in its segment (incl. itself)
    //   p_seg    its segment (0 .. nseg-1)
    int64_t p_it = 0, p_kt = 0, p_left = 0, p_seg = -1;
    int p_s = 0;
    uint32_t p_ph = 0;
    // first k-tile and length of segment j
    auto seg_first_kt = [&](int64_t j, int64_t& len) -> int64_t {
        if (j < full_rounds) { len = nkt; return kt0; }
        if (j == full_rounds) { len = lenA; return kA; }
        len = lenB; return 0;
    };
    auto produce_next = [&]() {
        if (p_left == 0) {                                        // enter the next segment
            ++p_seg;
            p_kt = seg_first_kt(p_seg, p_left);
        }
        const int64_t kt = p_kt;
        // k-tile of the iteration after this one (its coordinates ride along with this stage)
        int64_t ktn;
        if (p_left > 1) {
            ktn = kt + 1;
            if (p_seg < full_rounds && ktn == nkt) ktn = 0;
        } else if (p_seg + 1 < nseg) {
            int64_t len_unused;
            ktn = seg_first_kt(p_seg + 1, len_unused);
        } else {
            ktn = kt;
        }
        if (p.win_epochs > 0 && (p_it & (((int64_t)1 << p.epoch_shift) - 1)) == 0) {
            const int64_t e = p_it >> p.epoch_shift;
            atomicAdd(p.sync_cnt + e, 1u);
            const int64_t ew = e - p.win_epochs;
            if (window_on && ew >= 0) {
                const unsigned target = ((ew << p.epoch_shift) < full_it) ? gridDim.x : tail_ctas_beyond((ew << p.epoch_shift) - full_it);
                const volatile unsigned int* c = p.sync_cnt + ew;
                const long long t0 = clock64();
                while (*c < target) {
                    if (clock64() - t0 > KC_SPIN_LIMIT) { window_on = false; break; }
                }
            }
        }
        mbar_wait(&empty[p_s], p_ph ^ 1u);
        double* xs = smem + (size_t)p_s * stage_doubles;
        double* us = xs + KC_BK * ld;
        mbar_expect_tx(&full[p_s], stage_bytes);
        if (p.l2_hint) bulk_g2s_hint(xs, p.X + kt * KC_BK * p.ld, KC_BK * ld * 8, &full[p_s], xpolicy);
        else bulk_g2s(xs, p.X + kt * KC_BK * p.ld, KC_BK * ld * 8, &full[p_s]);
        if (KIND == KC_KIND_TABLE) {
            int* usi = reinterpret_cast<int*>(us);
            for (int k = 0; k < DIM; ++k)
                bulk_g2s(usi + k * KC_BK, p.lat + k * p.n_pad + ktn * KC_BK, KC_BK * 4, &full[p_s]);
        } else {
            for (int k = 0; k < DIM; ++k)
                bulk_g2s(us + k * KC_BK, p.u + k * p.n_pad + ktn * KC_BK, KC_BK * 8, &full[p_s]);
        }
        // advance
        ++p_it;
        --p_left;
        p_kt = kt + 1;
        if (p_seg < full_rounds && p_kt == nkt) p_kt = 0;
        if (++p_s == nstages) { p_s = 0; p_ph ^= 1u; }
    };
    if (tid == 0) {
        for (int64_t i = 0; i < lookahead && i < total_it; ++i) produce_next();
    }

    const int g = lane >> 2;      // fragment row (A, C) / column (B)
    const int t = lane & 3;       // fragment k index (A, B) / column pair (C)
    // generation role inside the row group: 128 threads x 4 entries = 16 rows x 32 points
    const int gtid = tid & 127;                  // thread index within the row group
    const int grow_in_tile = rg * 16 + (gtid >> 3);
    // my four points of the k-tile: columns gj0, gj0+1 and gj0+16, gj0+17 (gcol(e)).  The eight threads of a row
    // then publish 16-byte pairs at units 0..7 and 8..15 of the row: conflict-free 128-bit stores (round 1 used
    // four adjacent columns per thread: units 0,2,..,14 -> every quarter-warp store 2-way conflicted, 56 % of all
    // shared-store wavefronts of the kernel).
    const int gj0 = (gtid & 7) * 2;
    auto gcol = [&](int e) -> int { return gj0 + (e & 1) + ((e >> 1) << 4); };
    int li[DIM];
    auto row_coords = [&](int64_t base_rg, double (&ui)[DIM]) {
        int64_t gpt = p.row0 + base_rg * 16 + grow_in_tile;           // global point of my generated row
        if (gpt > p.n - 1) gpt = p.n - 1;                             // tail rows: clamp (never stored)
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            if (KIND == KC_KIND_TABLE) li[k] = p.lat[k * p.n_pad + gpt];
            else ui[k] = p.u[k * p.n_pad + gpt];
        }
    };
    // table variant: lattice offset of (my row, column point e of the tile) -> table index
    auto table_index = [&](const int* usi, int e) -> int {
        int idx = 0;
#pragma unroll
        for (int k = DIM - 1; k >= 0; --k) {
            int dk = li[k] - usi[k * KC_BK + gcol(e)];
            dk = dk < 0 ? -dk : dk;
            idx = (k == DIM - 1) ? dk : idx * (k == 0 ? p.nx : p.ny) + dk;
        }
        return idx;
    };
    auto gen4 = [&](const double (&ui)[DIM], const double* u0, int64_t ustride, double (&v)[4]) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            double r2 = (KIND == GSI_KERNEL_EXPONENTIAL) ? 1e-300 : 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const double dk = ui[k] - u0[k * ustride + gcol(e)];
                r2 = fma(dk, dk, r2);
            }
            v[e] = kern_eval<KIND>(r2, p.beta, etab);
        }
    };
    auto store4 = [&](double* as, const double (&v)[4]) {
        double2* dst = reinterpret_cast<double2*>(as + grow_in_tile * KC_AP + gj0);
        dst[0] = make_double2(v[0], v[1]);
        dst[8] = make_double2(v[2], v[3]);          // + 16 columns
    };

    int c_s = 0;                       // consumer pipeline stage / mbarrier phase of iteration `it`
    uint32_t c_ph = 0;
    double ui[DIM];
    if (nseg > 0) {
        // prologue: kernel values of (first segment, its first k-tile) straight from global coordinates
        const Seg g0 = seg_info(0);
        int64_t j0, ktf;
        locate(0, j0, ktf);
        row_coords(g0.base_rg, ui);
        if (rg < g0.nact) {
            double v[4];
            if (KIND == KC_KIND_TABLE) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int idx = 0;
#pragma unroll
                    for (int k = DIM - 1; k >= 0; --k) {
                        int dk = li[k] - p.lat[k * p.n_pad + ktf * KC_BK + gcol(e)];
                        dk = dk < 0 ? -dk : dk;
                        idx = (k == DIM - 1) ? dk : idx * (k == 0 ? p.nx : p.ny) + dk;
                    }
                    v[e] = __ldg(p.table + idx);
                }
            } else {
                gen4(ui, p.u + ktf * KC_BK, p.n_pad, v);
            }
            store4(a_tiles, v);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&abar[rg]);
    }
    // (loop state is kept small on purpose: the kernel sits at the 128-register limit and every value that
    //  lives across the k loop costs the generation / MMA interleave its scheduling freedom)
    int par = 0;                                                      // parity of the A-tile buffer of this iteration
    for (int64_t j = 0; j < nseg; ++j) {
        int klen;
        bool active;
        {
            const Seg sg = seg_info(j);
            klen = (int)sg.klen;
            active = rg < sg.nact;                                    // warp-uniform
        }
        double acc[2][NBW][2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) { acc[h][nb][0] = 0.0; acc[h][nb][1] = 0.0; }

        for (int ktl = 0; ktl < klen; ++ktl, par ^= 1) {
            if (tid == 0 && p_it < total_it) produce_next();           // (p_it == it + lookahead)
            const int s = c_s;
            mbar_wait(&full[s], c_ph);
            if (++c_s == nstages) { c_s = 0; c_ph ^= 1u; }
            __syncwarp();
            const double* xs = smem + (size_t)s * stage_doubles;
            const double* us = xs + KC_BK * ld;                       // coordinates of the next iteration's k-tile
            const double* as = a_tiles + (size_t)par * a_doubles;
            // ---- kernel values of the NEXT iteration's block (independent of the MMAs below: the two
            //      instruction streams overlap on the shared FP64 pipe)
            bool gen_next = active;
            if (ktl + 1 == klen) {                                    // the next block belongs to the next segment
                gen_next = false;
                if (j + 1 < nseg) {
                    const Seg sn = seg_info(j + 1);
                    row_coords(sn.base_rg, ui);
                    gen_next = rg < sn.nact;
                }
            }
            double* anext = a_tiles + (size_t)(par ^ 1) * a_doubles + grow_in_tile * KC_AP + gj0;
            double tv[4] = {0.0, 0.0, 0.0, 0.0};
            if (KIND == KC_KIND_TABLE && gen_next) {
                // structured grid: the four kernel values are table look-ups (L1/L2 hits, no FP64 pipe
                // work); issued now, consumed after the MMA loop
                const int* usi = reinterpret_cast<const int*>(us);
#pragma unroll
                for (int e = 0; e < 4; ++e) tv[e] = __ldg(p.table + table_index(usi, e));
            }
            // ---- tensor-core phase on the current k-tile: wait until the whole row group has
            //      published this block (and has stopped reading the other buffer)
            mbar_wait(&abar[rg], (uint32_t)par);
            __syncwarp();
            const double* arow0 = as + (rg * 16 + g) * KC_AP + t;
            const double* arow1 = arow0 + 8 * KC_AP;
#pragma unroll
            for (int ks = 0; ks < KC_BK / 4; ++ks) {
                // one kernel value of the NEXT block every second k-step: a single exp chain
                // is live at a time and its DFMAs interleave with this step's DMMAs
                if (KIND != KC_KIND_TABLE && (ks & 1) == 0 && gen_next) {
                    const int e = ks >> 1;
                    double r2 = (KIND == GSI_KERNEL_EXPONENTIAL) ? 1e-300 : 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const double dk = ui[k] - us[k * KC_BK + gcol(e)];
                        r2 = fma(dk, dk, r2);
                    }
                    anext[(e & 1) + ((e >> 1) << 4)] = kern_eval<KIND>(r2, p.beta, etab);
                }
                if (active) {
                    const double a0 = arow0[ks * 4];
                    const double a1 = arow1[ks * 4];
                    const double* xrow = xs + (ks * 4 + t) * ld + nb0 * 8 + g;
#pragma unroll
                    for (int nb = 0; nb < NBW; ++nb) {
                        if (nb0 + nb < NB) {
                            const double b = xrow[nb * 8];
                            dmma884(acc[0][nb][0], acc[0][nb][1], a0, b);
                            dmma884(acc[1][nb][0], acc[1][nb][1], a1, b);
                        }
                    }
                }
            }
            if (KIND == KC_KIND_TABLE && gen_next) {
                double2* dst = reinterpret_cast<double2*>(anext);
                dst[0] = make_double2(tv[0], tv[1]);
                dst[8] = make_double2(tv[2], tv[3]);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                mbar_arrive(&abar[rg]);           // next block published; current block no longer read
            }
        }

        const Seg sg = seg_info(j);
        if (sg.slot < 0) {
            // epilogue of a full round: W = sigma2 * acc + nugget * X[global row]
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t lrow = (sg.base_rg + rg) * 16 + h * 8 + g;
                if (active && lrow < p.mloc) {
                    double* wrow = p.W + lrow * p.ldw + nb0 * 8 + 2 * t;
                    const double* xg = p.X + (p.row0 + lrow) * p.ld + nb0 * 8 + 2 * t;
#pragma unroll
                    for (int nb = 0; nb < NBW; ++nb) {
                        if (nb0 + nb < NB) {
                            double2 v;
                            v.x = p.sigma2 * acc[h][nb][0];
                            v.y = p.sigma2 * acc[h][nb][1];
                            if (p.nugget != 0.0) {
                                const double2 xv = *reinterpret_cast<const double2*>(xg + nb * 8);
                                v.x += p.nugget * xv.x;
                                v.y += p.nugget * xv.y;
                            }
                            *reinterpret_cast<double2*>(wrow + nb * 8) = v;
                        }
                    }
                }
            }
        } else if (active) {
            // tail segment: raw partial sums of this k range -> slot (kcov_tail_fixup_kernel combines them)
            double* part = p.partial + (size_t)sg.slot * KC_BM * p.ldw;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double* prow = part + (size_t)(rg * 16 + h * 8 + g) * p.ldw + nb0 * 8 + 2 * t;
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb)
                    if (nb0 + nb < NB) *reinterpret_cast<double2*>(prow + nb * 8) = make_double2(acc[h][nb][0], acc[h][nb][1]);
            }
        }
    }
}

// Tail tiles: W[rows of tile t] = sigma2 * (sum of the k-range pieces in ascending CTA order) + nugget * X[rows].
// Piece of CTA b for tile t lives in slot 2b (tile t is the first tile of b's share) or 2b + 1.
// Grid: (tail tiles, KC_BM / KF_ROWS row chunks); the slot list of the tile is built once per CTA (the
// round-2 version divided 64-bit integers per piece and element on 9 CTAs: 141 us for 9 tiles at n = 10 000).
constexpr int KF_ROWS = 8;
constexpr int KF_MAX_PIECES = 160;     // pieces of one tile <= CTAs of the main grid (<= SMs, checked at launch)
__global__ void kcov_tail_fixup_kernel(KcovParams p, int64_t nkt, int64_t tail_rg0, int grid_main, int cols) {
    __shared__ int s_slot[KF_MAX_PIECES];
    const int64_t tile = blockIdx.x;
    const int64_t ulo = tile * nkt, uhi = ulo + nkt;                  // this tile's units
    const int b_first = (int)(ulo / p.tail_u);
    int b_last = (int)((uhi - 1) / p.tail_u);
    if (b_last > grid_main - 1) b_last = grid_main - 1;
    const int npieces = b_last - b_first + 1;
    for (int k = threadIdx.x; k < npieces; k += blockDim.x) {
        const int b = b_first + k;
        const int64_t first_tile = ((int64_t)b * p.tail_u) / nkt;
        s_slot[k] = 2 * b + (tile == first_tile ? 0 : 1);
    }
    __syncthreads();
    const int64_t row_base = tail_rg0 * 16 + tile * KC_BM;
    const int r_lo = blockIdx.y * KF_ROWS;
    for (int idx = threadIdx.x; idx < KF_ROWS * cols; idx += blockDim.x) {
        const int r = r_lo + idx / cols, c = idx % cols;
        const int64_t lrow = row_base + r;
        if (lrow >= p.mloc) continue;
        const double* src = p.partial + (size_t)r * p.ldw + c;
        double s = 0.0;
        for (int k = 0; k < npieces; ++k) s += src[(size_t)s_slot[k] * KC_BM * p.ldw];
        double v = p.sigma2 * s;
        if (p.nugget != 0.0) v += p.nugget * p.X[(p.row0 + lrow) * p.ld + c];
        p.W[lrow * p.ldw + c] = v;
    }
}

template <int NB, int KIND, int DIM>
static void launch_kcov(gsi_ctx* ctx, const KcovParams& p0) {
    KcovParams p = p0;
    const int ld = NB * 8 + 4;
    const size_t stage_bytes = (size_t)(KC_BK * ld + 3 * KC_BK) * sizeof(double);
    const size_t a_bytes = (size_t)2 * KC_BM * KC_AP * sizeof(double);
    int stages = (int)((224 * 1024 - a_bytes - 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + a_bytes + (2 * stages + 4) * sizeof(uint64_t) + 64 * sizeof(double);
    auto kfn = kcov_gemm_kernel<NB, KIND, DIM>;
    GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    GSI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, KC_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t total_rg = (p.mloc + 15) / 16;                 // 16-row groups; a CTA runs up to 4 at a time
    int64_t grid = (int64_t)ctx->num_sms * occ;
    if (grid > total_rg) grid = total_rg;
    if (grid < 1) grid = 1;
    p.pad0 = 0;
    p.win_epochs = ctx->kcov_window > 0 ? ctx->kcov_window : 0;
    p.epoch_shift = ctx->kcov_epoch_shift;
    p.sync_cnt = nullptr;
    if (p.win_epochs > 0) {
        // one arrival counter per epoch of the longest CTA schedule (full rounds + tail round)
        const int64_t nkt = (p.n + KC_BK - 1) / KC_BK;
        const int64_t rounds = total_rg / (grid * 4) + (total_rg % (grid * 4) != 0 ? 1 : 0);
        const size_t n_epochs = (size_t)((rounds * nkt) >> p.epoch_shift) + 2;
        if (n_epochs > ctx->sweep_cnt_n) {
            if (ctx->sweep_cnt) GSI_CUDA(cudaFree(ctx->sweep_cnt));
            ctx->sweep_cnt = nullptr; ctx->sweep_cnt_n = 0;
            GSI_CUDA(cudaMalloc(&ctx->sweep_cnt, n_epochs * sizeof(unsigned int)));
            ctx->sweep_cnt_n = n_epochs;
        }
        GSI_CUDA(cudaMemsetAsync(ctx->sweep_cnt, 0, n_epochs * sizeof(unsigned int), ctx->stream));
        p.sync_cnt = ctx->sweep_cnt;
    }
    // stream-K tail: tiles left over after the full rounds, cut into equal k shares (one per CTA)
    const int64_t nkt_h = (p.n + KC_BK - 1) / KC_BK;
    const int64_t full_rounds_h = total_rg / (grid * 4);
    const int64_t tail_rg0_h = full_rounds_h * grid * 4;
    const int64_t tail_tiles_h = (total_rg - tail_rg0_h + 3) / 4;
    p.tail_u = tail_tiles_h > 0 ? (tail_tiles_h * nkt_h + grid - 1) / grid : 0;
    p.partial = nullptr;
    size_t part_bytes = 0;
    if (tail_tiles_h > 0) {
        part_bytes = (size_t)2 * grid * KC_BM * p.ldw * sizeof(double);
        p.partial = static_cast<double*>(pool_alloc(ctx, part_bytes));
    }
    struct PartGuard { gsi_ctx* c; void* q; size_t b; ~PartGuard() { if (q) pool_free(c, q, b); } } pguard{ctx, p.partial, part_bytes};
    kfn<<<(unsigned)grid, KC_THREADS, smem, ctx->stream>>>(p);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
    if (tail_tiles_h > 0) {
        GSI_REQUIRE(grid <= KF_MAX_PIECES, GSI_ERR_UNSUPPORTED, "kernelcov apply: more CTAs than the tail fix-up handles");
        kcov_tail_fixup_kernel<<<dim3((unsigned)tail_tiles_h, KC_BM / KF_ROWS), 256, 0, ctx->stream>>>(p, nkt_h, tail_rg0_h,
                                                                                                      (int)grid, 8 * NB);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
}

template <int NB, int KIND>
static void dispatch_dim(gsi_ctx* ctx, const KcovParams& p, int dim) {
    if (dim <= 2) launch_kcov<NB, KIND, 2>(ctx, p);
    else launch_kcov<NB, KIND, 3>(ctx, p);
}

template <int NB>
static void dispatch_kind(gsi_ctx* ctx, const KcovParams& p, int kind, int dim) {
    switch (kind) {
        case GSI_KERNEL_EXPONENTIAL: dispatch_dim<NB, GSI_KERNEL_EXPONENTIAL>(ctx, p, dim); break;
        case GSI_KERNEL_GAUSSIAN: dispatch_dim<NB, GSI_KERNEL_GAUSSIAN>(ctx, p, dim); break;
        case GSI_KERNEL_POWERLAW: dispatch_dim<NB, GSI_KERNEL_POWERLAW>(ctx, p, dim); break;
        case KC_KIND_TABLE: dispatch_dim<NB, KC_KIND_TABLE>(ctx, p, dim); break;
        default: throw Error(GSI_ERR_INVALID_ARGUMENT, "unknown covariance kernel kind");
    }
}

void kcov_apply(gsi_op* op, const gsi_buf* X, gsi_buf* W) {
    gsi_ctx* ctx = op->ctx;
    GSI_REQUIRE(X->layout == GSI_LAYOUT_TALL && W->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT,
                "kernelcov apply needs TALL buffers");
    GSI_REQUIRE(X->rows == op->n, GSI_ERR_DIMENSION_MISMATCH, "kernelcov apply: X must have n rows");
    GSI_REQUIRE(W->rows == op->mloc, GSI_ERR_DIMENSION_MISMATCH, "kernelcov apply: W must have mloc rows");
    GSI_REQUIRE(X->cols == W->cols, GSI_ERR_DIMENSION_MISMATCH, "kernelcov apply: X/W column mismatch");
    if (X->cols > kMaxCols) {
        // wide iterate (K + p > 256): the product runs on 256-column chunks packed into compact buffers
        for (int64_t c0 = 0; c0 < X->cols; c0 += kMaxCols) {
            const int64_t w = (X->cols - c0 < kMaxCols) ? X->cols - c0 : kMaxCols;
            BufPtr xc = make_buf(ctx, GSI_LAYOUT_TALL, X->rows, w);
            BufPtr wc = make_buf(ctx, GSI_LAYOUT_TALL, W->rows, w);
            tall_cols_copy(ctx, X->d + c0, X->ld, xc->d, xc->ld, X->rows, w);
            kcov_apply(op, xc.get(), wc.get());
            tall_cols_copy(ctx, wc->d, wc->ld, W->d + c0, W->ld, W->rows, w);
        }
        return;
    }
    const int nb = nb_for_cols(X->cols);
    GSI_REQUIRE(X->ld == 8 * nb + 4 && W->ld == X->ld, GSI_ERR_INVALID_ARGUMENT, "kernelcov apply: bad pitch");
    KcovParams p;
    p.X = X->d; p.W = W->d; p.u = op->ucoords;
    p.lat = op->lattice; p.table = op->table; p.nx = op->grid_nx; p.ny = op->grid_ny;
    p.sweep_groups = ctx->kcov_sweep_groups; p.sweep_div = ctx->kcov_sweep_div; p.l2_hint = ctx->kcov_l2_hint;
    p.n = op->n; p.n_pad = op->n_pad; p.row0 = op->row0; p.mloc = op->mloc;
    p.ld = X->ld; p.ldw = W->ld;
    p.sigma2 = op->sigma2; p.nugget = op->nugget; p.beta = op->beta;
    p.stages = 0;
    switch (nb) {
#define GSI_CASE(N) case N: dispatch_kind<N>(ctx, p, op->table ? KC_KIND_TABLE : op->kind, op->dim); break;
        GSI_NB_LIST(GSI_CASE)
#undef GSI_CASE
        default: throw Error(GSI_ERR_UNSUPPORTED, "kernelcov apply: unsupported column-block count");
    }
}

}  // namespace gsi
