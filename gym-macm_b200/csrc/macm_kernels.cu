// macm_kernels.cu -- hand-written sm_100a kernels for the batched gym-macm step.
//
// One kernel launch == one Flock.step / TDM.step (mvmnt.py:81-140, combat.py:104-184) for every
// environment of the batch, i.e. action decode + b2World::Step + reward pass + observation pass.
//
// Mapping.  An environment is owned by a GROUP of G lanes of one warp (G = 4, 8, 16 or 32); each
// lane owns APL agents (agent i lives in lane i % G, slot i / G), so N <= G*APL <= 64.  Groups
// never talk to each other, so all synchronisation is __syncwarp / ballot on the group mask and
// a block is just a bag of independent groups (no __syncthreads anywhere).  The bodies, the fat
// AABBs, the contact adjacency (one 64-bit row per agent) and the touching contacts of an
// environment are staged in shared memory; HBM is touched once per array per step with
// coalesced float4 / float2 accesses.
//
// Exactness.  The engine arithmetic is Box2D's: fp32, round-to-nearest, NO fused multiply-add
// (this file is compiled with -fmad=false, IEEE division and square root), evaluated in the
// same association order as Box2D 2.3.0.  Contacts are solved in Box2D's island order: a
// depth-first traversal from the last-created body over contact edges newest-first.  Two
// contacts that share no body commute bit-exactly, so the ordered list is level-scheduled
// (level = 1 + max(level of the previous contact of either body)) and each level runs across
// lanes; contacts of the same level touch disjoint bodies.
//
// The reference's host arithmetic is Python float64 (angle update, sin/cos, force); that part is
// done in fp64 here and rounded to fp32 exactly where pybox2d's SWIG layer rounds.
#include <float.h>

#include "macm_sim.h"

namespace {

// ------------------------------------------------------------------------------------------
// shared-memory layout of one environment
// ------------------------------------------------------------------------------------------
template <int NC>
struct Lay {
    static constexpr int PX = 0, PY = PX + 4 * NC, VX = PY + 4 * NC, VY = VX + 4 * NC;
    static constexpr int FLX = VY + 4 * NC, FLY = FLX + 4 * NC, FHX = FLY + 4 * NC, FHY = FHX + 4 * NC;
    static constexpr int ADJL = FHY + 4 * NC, ADJH = ADJL + 4 * NC, NEWL = ADJH + 4 * NC, NEWH = NEWL + 4 * NC;
    static constexpr int ISLMIN = NEWH + 4 * NC;
    static constexpr int LABEL = ISLMIN + 4 * NC, STACK = LABEL + NC, LASTLVL = STACK + NC;
    static constexpr int ISLACT = LASTLVL + NC, ISLBAD = ISLACT + NC, HEAD = ISLBAD + NC;
    static constexpr int MISC = (HEAD + NC + 15) / 16 * 16;  // 4 x u32
    static constexpr int FIXED = MISC + 16;
    // per touching contact (capacity TC, a multiple of 16):
    //   float nx, ny, nI, tI ; u16 slot, ord ; u8 a, b, lvl, nxt_a, nxt_b, taken      = 26 bytes
    __host__ __device__ static constexpr int bytes(int TC) { return (FIXED + 26 * TC + 15) / 16 * 16; }
};

struct EnvS {
    unsigned char* base;
    int TC;
    template <int NC> __device__ float* px() const { return (float*)(base + Lay<NC>::PX); }
    template <int NC> __device__ float* py() const { return (float*)(base + Lay<NC>::PY); }
    template <int NC> __device__ float* vx() const { return (float*)(base + Lay<NC>::VX); }
    template <int NC> __device__ float* vy() const { return (float*)(base + Lay<NC>::VY); }
    template <int NC> __device__ float* flx() const { return (float*)(base + Lay<NC>::FLX); }
    template <int NC> __device__ float* fly() const { return (float*)(base + Lay<NC>::FLY); }
    template <int NC> __device__ float* fhx() const { return (float*)(base + Lay<NC>::FHX); }
    template <int NC> __device__ float* fhy() const { return (float*)(base + Lay<NC>::FHY); }
    template <int NC> __device__ uint32_t* adj_lo() const { return (uint32_t*)(base + Lay<NC>::ADJL); }
    template <int NC> __device__ uint32_t* adj_hi() const { return (uint32_t*)(base + Lay<NC>::ADJH); }
    template <int NC> __device__ uint32_t* new_lo() const { return (uint32_t*)(base + Lay<NC>::NEWL); }
    template <int NC> __device__ uint32_t* new_hi() const { return (uint32_t*)(base + Lay<NC>::NEWH); }
    template <int NC> __device__ uint32_t* isl_min() const { return (uint32_t*)(base + Lay<NC>::ISLMIN); }
    template <int NC> __device__ uint8_t* label() const { return base + Lay<NC>::LABEL; }
    template <int NC> __device__ uint8_t* stack() const { return base + Lay<NC>::STACK; }
    template <int NC> __device__ uint8_t* lastlvl() const { return base + Lay<NC>::LASTLVL; }
    template <int NC> __device__ uint8_t* isl_act() const { return base + Lay<NC>::ISLACT; }
    template <int NC> __device__ uint8_t* isl_bad() const { return base + Lay<NC>::ISLBAD; }
    template <int NC> __device__ uint32_t* misc() const { return (uint32_t*)(base + Lay<NC>::MISC); }
    template <int NC> __device__ float* t_nx() const { return (float*)(base + Lay<NC>::FIXED); }
    template <int NC> __device__ float* t_ny() const { return t_nx<NC>() + TC; }
    template <int NC> __device__ float* t_nI() const { return t_nx<NC>() + 2 * TC; }
    template <int NC> __device__ float* t_tI() const { return t_nx<NC>() + 3 * TC; }
    template <int NC> __device__ uint16_t* t_slot() const { return (uint16_t*)(t_nx<NC>() + 4 * TC); }
    template <int NC> __device__ uint16_t* ord() const { return t_slot<NC>() + TC; }
    template <int NC> __device__ uint8_t* t_a() const { return (uint8_t*)(ord<NC>() + TC); }
    template <int NC> __device__ uint8_t* t_b() const { return t_a<NC>() + TC; }
    template <int NC> __device__ uint8_t* lvl() const { return t_b<NC>() + TC; }
    template <int NC> __device__ uint8_t* nxt_a() const { return t_b<NC>() + 2 * TC; }
    template <int NC> __device__ uint8_t* nxt_b() const { return t_b<NC>() + 3 * TC; }
    template <int NC> __device__ uint8_t* taken() const { return t_b<NC>() + 4 * TC; }
    template <int NC> __device__ uint8_t* head() const { return base + Lay<NC>::HEAD; }
};

// ------------------------------------------------------------------------------------------
// group-of-G-lanes primitives
// ------------------------------------------------------------------------------------------
template <int G>
struct Grp {
    unsigned mask;
    int shift, gl;
    __device__ Grp()
    {
        const int lane = threadIdx.x & 31;
        gl = lane % G;
        shift = lane - gl;
        mask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << shift);
    }
    __device__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) & mask) >> shift; }
    __device__ void sync() const { __syncwarp(mask); }
    __device__ unsigned below() const { return (1u << gl) - 1u; }
    __device__ unsigned reduce_or(unsigned v) const { return __reduce_or_sync(mask, v); }
    __device__ unsigned reduce_min(unsigned v) const { return __reduce_min_sync(mask, v); }
    __device__ unsigned reduce_max(unsigned v) const { return __reduce_max_sync(mask, v); }
    __device__ int shfl(int v, int src) const { return __shfl_sync(mask, v, shift + src); }
    // inclusive prefix sum over the group lanes
    __device__ int scan_incl(int v) const
    {
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            int o = __shfl_up_sync(mask, v, d, G);
            if (gl >= d) v += o;
        }
        return v;
    }
};

// b2TestOverlap(b2AABB, b2AABB): d1 = b.lo - a.hi, d2 = a.lo - b.hi; overlap iff no component > 0.
// With IEEE gradual underflow (no -ftz) x - y > 0 <=> x > y, so the subtractions are not needed.
__device__ __forceinline__ bool aabb_overlap(float alx, float aly, float ahx, float ahy, float blx, float bly,
                                             float bhx, float bhy)
{
    return !(blx > ahx || bly > ahy || alx > bhx || aly > bhy);
}

__device__ __forceinline__ float b2min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float b2max(float a, float b) { return a > b ? a : b; }
__device__ __forceinline__ float b2clamp(float a, float lo, float hi) { return b2max(lo, b2min(a, hi)); }

// b2Vec2::Normalize (unchanged when shorter than b2_epsilon)
__device__ __forceinline__ void b2normalize(float& x, float& y)
{
    float len = sqrtf(x * x + y * y);
    if (len < B2_EPSILON) return;
    float inv = 1.0f / len;
    x *= inv;
    y *= inv;
}

// t - sign(t)*2*pi if |t| > pi else t (mvmnt.py:199), in fp32 for the observation outputs
__device__ __forceinline__ float wrap_pi_f(float t)
{
    const float PI_F = 3.14159265358979f, TWO_PI_F = 6.28318530717959f;
    if (fabsf(t) > PI_F) t = t - copysignf(TWO_PI_F, t);
    return t;
}

// One velocity-constraint pass of one contact (b2ContactSolver::SolveVelocityConstraints,
// pointCount == 1, invI = 0): tangent (friction) first, then normal.
__device__ __forceinline__ void solve_velocity(float nx, float ny, float friction, float mass_n, float mass_t,
                                               float inv_mass, float& nI, float& tI, float& vax, float& vay,
                                               float& vbx, float& vby)
{
    const float tx = ny, ty = -nx;  // b2Cross(normal, 1.0f)
    {
        float dvx = vbx - vax, dvy = vby - vay;
        float vt = dvx * tx + dvy * ty;
        float lambda = mass_t * (-vt);
        float maxf = friction * nI;
        float ni = b2clamp(tI + lambda, -maxf, maxf);
        lambda = ni - tI;
        tI = ni;
        float Px = lambda * tx, Py = lambda * ty;
        vax -= inv_mass * Px; vay -= inv_mass * Py;
        vbx += inv_mass * Px; vby += inv_mass * Py;
    }
    {
        float dvx = vbx - vax, dvy = vby - vay;
        float vn = dvx * nx + dvy * ny;
        float lambda = -mass_n * vn;
        float ni = b2max(nI + lambda, 0.0f);
        lambda = ni - nI;
        nI = ni;
        float Px = lambda * nx, Py = lambda * ny;
        vax -= inv_mass * Px; vay -= inv_mass * Py;
        vbx += inv_mass * Px; vby += inv_mass * Py;
    }
}

// One position-constraint pass of one contact (b2ContactSolver::SolvePositionConstraints +
// b2PositionSolverManifold::Initialize, e_circles).  Returns the separation it saw.
__device__ __forceinline__ float solve_position(float radius, float k_sum, float inv_mass, float& cax, float& cay,
                                                float& cbx, float& cby)
{
    float nx = cbx - cax, ny = cby - cay;
    const float dx = nx, dy = ny;
    b2normalize(nx, ny);
    float sep = (dx * nx + dy * ny) - radius - radius;
    float C = b2clamp(B2_BAUMGARTE * (sep + B2_LINEAR_SLOP), -B2_MAX_LINEAR_CORRECTION, 0.0f);
    float imp = k_sum > 0.0f ? -C / k_sum : 0.0f;
    float Px = imp * nx, Py = imp * ny;
    cax -= inv_mass * Px; cay -= inv_mass * Py;
    cbx += inv_mass * Px; cby += inv_mass * Py;
    return sep;
}

__device__ __forceinline__ void set_bit64(uint32_t* lo, uint32_t* hi, int row, int bit)
{
    if (bit < 32) atomicOr(&lo[row], 1u << bit);
    else atomicOr(&hi[row], 1u << (bit - 32));
}

// ------------------------------------------------------------------------------------------
// b2ContactManager::FindNewContacts for one environment (group-cooperative).
//   moved: 64-bit mask of proxies in the broadphase move buffer.
//   Every other proxy whose fat AABB overlaps a moved one forms a pair; pairs are sorted
//   lexicographically, de-duplicated, and those without a contact yet are created in that
//   order (lower index = fixtureA).  Here: row i of a bit matrix holds the partners j > i;
//   emitting rows in index order and bits in ascending order IS the sorted order.
// Appends to the HBM contact list at `cnt`; returns the new count (uniform over the group).
// ------------------------------------------------------------------------------------------
template <int G, int APL>
__device__ int find_new_contacts(const Grp<G>& g, const EnvS& S, const SimConst& P, uint64_t moved, uint64_t alive,
                                 int cnt, uint32_t* c_ab, float2* c_imp, bool& overflow)
{
    constexpr int NC = G * APL;
    float* flx = S.flx<NC>(); float* fly = S.fly<NC>(); float* fhx = S.fhx<NC>(); float* fhy = S.fhy<NC>();
    uint32_t* adj_lo = S.adj_lo<NC>(); uint32_t* adj_hi = S.adj_hi<NC>();
    uint32_t* new_lo = S.new_lo<NC>(); uint32_t* new_hi = S.new_hi<NC>();

    uint64_t hit[APL];
    float olx[APL], oly[APL], ohx[APL], ohy[APL];
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        hit[s] = 0;
        olx[s] = flx[i]; oly[s] = fly[i]; ohx[s] = fhx[i]; ohy[s] = fhy[i];
        new_lo[i] = 0; new_hi[i] = 0;
    }
    g.sync();
    for (uint64_t mm = moved; mm; mm &= mm - 1) {
        const int m = __ffsll((long long)mm) - 1;
        const float mlx = flx[m], mly = fly[m], mhx = fhx[m], mhy = fhy[m];
#pragma unroll
        for (int s = 0; s < APL; ++s) {
            const int i = g.gl + s * G;
            if (i != m && ((alive >> i) & 1) && aabb_overlap(mlx, mly, mhx, mhy, olx[s], oly[s], ohx[s], ohy[s]))
                hit[s] |= 1ull << m;
        }
    }
    // pair (m, i) with m < i belongs to row m: the owner of m finds it itself iff i moved too;
    // otherwise the owner of i posts it
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        if (!((moved >> i) & 1)) {
            uint64_t low = hit[s] & ((1ull << i) - 1ull);
            for (; low; low &= low - 1) {
                const int m = __ffsll((long long)low) - 1;
                set_bit64(new_lo, new_hi, m, i);
            }
        }
    }
    g.sync();
    uint64_t fresh[APL];
    int base = cnt;
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        uint64_t row = (hit[s] & ~((2ull << i) - 1ull)) | ((uint64_t)new_lo[i] | ((uint64_t)new_hi[i] << 32));
        uint64_t have = (uint64_t)adj_lo[i] | ((uint64_t)adj_hi[i] << 32);
        fresh[s] = row & ~have;
        const int c = __popcll(fresh[s]);
        const int incl = g.scan_incl(c);
        const int total = g.shfl(incl, G - 1);
        int pos = base + incl - c;
        for (uint64_t f = fresh[s]; f; f &= f - 1) {
            const int j = __ffsll((long long)f) - 1;
            if (pos < P.C) {
                c_ab[pos] = (uint32_t)i | ((uint32_t)j << 8);
                c_imp[pos] = make_float2(0.0f, 0.0f);
                set_bit64(adj_lo, adj_hi, i, j);
                set_bit64(adj_lo, adj_hi, j, i);
            } else {
                overflow = true;
            }
            ++pos;
        }
        base += total;
    }
    g.sync();
    return base < P.C ? base : P.C;
}

// ------------------------------------------------------------------------------------------
// observation pass for the agents a lane owns (Flock.get_obs, mvmnt.py:181-222)
// ------------------------------------------------------------------------------------------
template <int G, int APL>
__device__ void flock_observe(const Grp<G>& g, const EnvS& S, const SimConst& P, int env, const float* ang)
{
    constexpr int NC = G * APL;
    const float* px = S.px<NC>(); const float* py = S.py<NC>();
    const int N = P.N;
    float ox[APL], oy[APL], best[APL];
    int bi[APL];
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        ox[s] = px[i]; oy[s] = py[i];
        best[s] = __int_as_float(0x7f800000);
        bi[s] = -1;
    }
    // nearest other agent: strict '<' over ascending j keeps the lowest index on ties (mvmnt.py:194)
#pragma unroll 4
    for (int j = 0; j < N; ++j) {
        const float qx = px[j], qy = py[j];
#pragma unroll
        for (int s = 0; s < APL; ++s) {
            const float dx = qx - ox[s], dy = qy - oy[s];
            const float d2 = dx * dx + dy * dy;  // b2DistanceSquared, no FMA
            const bool take = (d2 < best[s]) && (j != g.gl + s * G);
            best[s] = take ? d2 : best[s];
            bi[s] = take ? j : bi[s];
        }
    }
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        if (i >= N) continue;
        const size_t gi = (size_t)env * N + i;
        float nn_d = __int_as_float(0x7f800000), nn_t = 0.0f;
        if (bi[s] >= 0) {
            nn_d = sqrtf(best[s]);
            nn_t = wrap_pi_f(atan2f(py[bi[s]] - oy[s], px[bi[s]] - ox[s]) - ang[s]);
        }
        const float2 tg = P.targets[(size_t)env * P.T + P.target_idx[i]];
        const float tdx = tg.x - ox[s], tdy = tg.y - oy[s];
        const float tg_r = sqrtf(tdx * tdx + tdy * tdy);
        const float tg_t = wrap_pi_f(atan2f(tdy, tdx) - ang[s]);
        P.nn_idx[gi] = bi[s];
        if (P.coord == MACM_COORD_POLAR) {
            reinterpret_cast<float4*>(P.obs)[gi] = make_float4(nn_d, nn_t, tg_r, tg_t);
        } else {
            float sn, cn, st, ct;
            sincosf(nn_t, &sn, &cn);
            sincosf(tg_t, &st, &ct);
            float2* o = reinterpret_cast<float2*>(P.obs) + gi * 3;
            o[0] = make_float2(nn_d, cn);
            o[1] = make_float2(sn, tg_r);
            o[2] = make_float2(ct, st);
        }
    }
}

// ------------------------------------------------------------------------------------------
// the step kernel
// ------------------------------------------------------------------------------------------
template <int G, int APL>
__global__ void __launch_bounds__(128) macm_flock_step_kernel(const __grid_constant__ SimConst P,
                                                              const void* __restrict__ actions)
{
    constexpr int NC = G * APL;
    constexpr int GPW = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp<G> g;
    const int warp = threadIdx.x >> 5;
    const int gidx = (threadIdx.x & 31) / G;
    const int slot_in_block = warp * GPW + gidx;
    const int env = blockIdx.x * (blockDim.x / 32) * GPW + slot_in_block;
    if (env >= P.E) return;  // whole group leaves together
    const int N = P.N;
    EnvS S;
    S.TC = P.TC;
    S.base = smem_raw + (size_t)slot_in_block * Lay<NC>::bytes(P.TC);

    float* px = S.px<NC>(); float* py = S.py<NC>(); float* vx = S.vx<NC>(); float* vy = S.vy<NC>();
    float* flx = S.flx<NC>(); float* fly = S.fly<NC>(); float* fhx = S.fhx<NC>(); float* fhy = S.fhy<NC>();
    uint32_t* adj_lo = S.adj_lo<NC>(); uint32_t* adj_hi = S.adj_hi<NC>();
    uint8_t* label = S.label<NC>();
    uint32_t* misc = S.misc<NC>();

    uint32_t* c_ab = P.c_ab + (size_t)env * P.C;
    float2* c_imp = P.c_imp + (size_t)env * P.C;

    // ---- phase 0: load state ---------------------------------------------------------------
    float cx[APL], cy[APL], wx[APL], wy[APL], ang[APL], slp[APL], Fx[APL], Fy[APL];
    float4 fatr[APL];
    bool valid[APL];
    uint64_t alive = 0;
    const int4 es = P.env_state[env];
    int cnt = P.c_cnt[env];
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        valid[s] = i < N;
        const size_t gi = (size_t)env * N + (valid[s] ? i : 0);
        const float4 pv = P.posvel[gi];
        const float2 as = P.angsleep[gi];
        fatr[s] = P.fat[gi];
        cx[s] = pv.x; cy[s] = pv.y; wx[s] = pv.z; wy[s] = pv.w; ang[s] = as.x; slp[s] = as.y;
        if (!valid[s]) {  // padding agents: parked far away, never alive
            cx[s] = cy[s] = 3.0e30f; wx[s] = wy[s] = 0.0f;
            fatr[s] = make_float4(3.0e30f, 3.0e30f, 3.0e30f, 3.0e30f);
        }
        px[i] = cx[s]; py[i] = cy[s];
        flx[i] = fatr[s].x; fly[i] = fatr[s].y; fhx[i] = fatr[s].z; fhy[i] = fatr[s].w;
        adj_lo[i] = 0; adj_hi[i] = 0;
        alive |= (uint64_t)g.ballot(valid[s]) << (s * G);
    }
    if (g.gl == 0) { misc[0] = 0; misc[1] = 0; misc[2] = 0; }

    // ---- phase 1: actions -> angle, force (mvmnt.py:97-129), float64 like the reference -----
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        Fx[s] = 0.0f; Fy[s] = 0.0f;
        if (!valid[s]) continue;
        const size_t gi = (size_t)env * N + i;
        if (P.action_mode == MACM_ACTION_DISCRETE) {
            const uint32_t act = reinterpret_cast<const uint32_t*>(actions)[gi];
            const int a0 = (int)(act & 0xff) - 1, a1 = (int)((act >> 8) & 0xff) - 1, a2 = (int)((act >> 16) & 0xff) - 1;
            // body.angle = body.angle + (a2-1) * rotation_speed * (1/hz)   -> SetTransform rounds to fp32
            float af = (float)((double)ang[s] + (double)a2 * P.rot_step);
            if (fabs((double)af) > NP_PI) {
                const double a = (double)af;
                const double sg = (a > 0.0) ? 1.0 : -1.0;
                af = (float)(a - sg * (2 * NP_PI));
            }
            ang[s] = af;
            const double A = (double)af;
            double s1, c1, s2, c2;
            sincos(A, &s1, &c1);
            sincos(A + NP_PI / 2, &s2, &c2);
            const double c = (a0 != 0 && a1 != 0) ? P.diag : 1.0;
            Fx[s] = (float)((c1 * (double)a0 + c2 * (double)a1) * c * P.force);
            Fy[s] = (float)((s1 * (double)a0 + s2 * (double)a1) * c * P.force);
        } else {
            const float2 ac = reinterpret_cast<const float2*>(actions)[gi];
            double x = (double)ac.x, y = (double)ac.y;
            if ((x * x + y * y) > 1) {  // bug-compatible with mvmnt.py:124-126
                x = sqrt(x * x / (x * x + y * y));
                y = sqrt(y * y / (x * x + y * y));
            }
            Fx[s] = (float)(x * P.force);
            Fy[s] = (float)(y * P.force);
        }
    }
    g.sync();

    bool overflow_c = false, overflow_t = false;

    // ---- phase 2a: new fixtures -> FindNewContacts before Collide (b2World::Step prologue) ----
    if (es.y & MACM_ENV_FRESH) {
        for (int base = 0; base < cnt; base += G) {
            const int k = base + g.gl;
            if (k < cnt) {
                const uint32_t ab = c_ab[k];
                set_bit64(adj_lo, adj_hi, ab & 0xff, (ab >> 8) & 0xff);
                set_bit64(adj_lo, adj_hi, (ab >> 8) & 0xff, ab & 0xff);
            }
        }
        g.sync();
        cnt = find_new_contacts<G, APL>(g, S, P, alive, alive, cnt, c_ab, c_imp, overflow_c);
#pragma unroll
        for (int s = 0; s < APL; ++s) { adj_lo[g.gl + s * G] = 0; adj_hi[g.gl + s * G] = 0; }
        g.sync();
    }

    // ---- phase 2: b2ContactManager::Collide --------------------------------------------------
    // destroy contacts whose fat AABBs stopped overlapping, narrowphase the rest, compact in
    // place (birth order is preserved), stage the touching ones for the solver
    int tc = 0;
    {
        uint8_t* t_a = S.t_a<NC>(); uint8_t* t_b = S.t_b<NC>();
        float* t_nI = S.t_nI<NC>(); float* t_tI = S.t_tI<NC>();
        uint16_t* t_slot = S.t_slot<NC>();
        int w = 0;
        bool dup = false;
        for (int base = 0; base < cnt; base += G) {
            const int k = base + g.gl;
            const bool in = k < cnt;
            uint32_t ab = 0;
            float2 imp = make_float2(0.0f, 0.0f);
            if (in) { ab = c_ab[k]; imp = c_imp[k]; }
            g.sync();  // every lane holds its record before any lane compacts over it
            const int a = ab & 0xff, b = (ab >> 8) & 0xff;
            const bool keep = in && aabb_overlap(flx[a], fly[a], fhx[a], fhy[a], flx[b], fly[b], fhx[b], fhy[b]);
            bool touch = false;
            if (keep) {
                const float dx = px[b] - px[a], dy = py[b] - py[a];
                touch = !((dx * dx + dy * dy) > P.rsum2);
            }
            const unsigned km = g.ballot(keep), tm = g.ballot(touch);
            if (keep) {
                const int pos = w + __popc(km & g.below());
                const bool was = (ab >> 16) & 1;
                const float nI = (touch && was) ? imp.x : 0.0f, tI = (touch && was) ? imp.y : 0.0f;
                c_ab[pos] = (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)touch << 16);
                const int tp = tc + __popc(tm & g.below());
                // touching contacts get their impulses from StoreImpulses after the solver
                if (!touch || tp >= P.TC) c_imp[pos] = make_float2(nI, tI);
                set_bit64(adj_lo, adj_hi, a, b);
                set_bit64(adj_lo, adj_hi, b, a);
                if (touch) {
                    if (tp < P.TC) {
                        t_a[tp] = (uint8_t)a; t_b[tp] = (uint8_t)b; t_nI[tp] = nI; t_tI[tp] = tI;
                        t_slot[tp] = (uint16_t)pos;
                        // does any body carry two touching contacts?
                        const uint32_t oa = atomicOr(&misc[a >> 5], 1u << (a & 31));
                        const uint32_t ob = atomicOr(&misc[b >> 5], 1u << (b & 31));
                        dup |= ((oa >> (a & 31)) & 1) | ((ob >> (b & 31)) & 1);
                    }
                }
            }
            w += __popc(km);
            tc += __popc(tm);
        }
        cnt = w;
        if (tc > P.TC) { overflow_t = true; tc = P.TC; }
        if (g.ballot(dup)) { if (g.gl == 0) misc[2] = 1; }
    }

    // ---- phase 3: integrate velocities (b2Island::Solve, first loop) -------------------------
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        // v += h * (gravityScale * gravity + invMass * force); v *= damping
        wx[s] += P.h * (P.inv_mass * Fx[s]);
        wy[s] += P.h * (P.inv_mass * Fy[s]);
        wx[s] *= P.damp;
        wy[s] *= P.damp;
        vx[i] = wx[s]; vy[i] = wy[s];
        label[i] = (uint8_t)i;
        S.isl_act<NC>()[i] = 1;
        S.isl_bad<NC>()[i] = 0;
        S.lastlvl<NC>()[i] = 0;
        S.head<NC>()[i] = 0xff;
    }
    g.sync();

    // ---- phase 4: islands (b2World::Solve DFS) -> solve order + levels -------------------------
    // Box2D seeds islands from the body list (last-created body first), pops a stack, and walks
    // each body's contact edges newest-first; contacts are solved in the order they are added.
    // One lane replays exactly that over per-body edge lists (head-inserted in birth order, like
    // b2ContactManager::AddPair does) and assigns each contact its level on the fly.
    int nlev = tc > 0 ? 1 : 0;
    const bool multi = misc[2] != 0;
    if (tc > 0) {
        uint8_t* t_a = S.t_a<NC>(); uint8_t* t_b = S.t_b<NC>();
        uint16_t* ord = S.ord<NC>(); uint8_t* lvl = S.lvl<NC>();
        if (!multi) {
            // every body has at most one touching contact: contacts are independent, any order
            // gives the same bits; islands are the pairs themselves (seed = higher index)
            for (int k = g.gl; k < tc; k += G) {
                ord[k] = (uint16_t)k;
                lvl[k] = 1;
                const int a = t_a[k], b = t_b[k];
                label[a] = (uint8_t)b;
                label[b] = (uint8_t)b;
            }
        } else {
            int L = 1;
            if (g.gl == 0) {
                uint8_t* stack = S.stack<NC>(); uint8_t* head = S.head<NC>(); uint8_t* lastlvl = S.lastlvl<NC>();
                uint8_t* nxt_a = S.nxt_a<NC>(); uint8_t* nxt_b = S.nxt_b<NC>(); uint8_t* taken = S.taken<NC>();
                for (int t = 0; t < tc; ++t) {
                    const int a = t_a[t], b = t_b[t];
                    nxt_a[t] = head[a]; head[a] = (uint8_t)t;
                    nxt_b[t] = head[b]; head[b] = (uint8_t)t;
                    taken[t] = 0;
                }
                // bodies with a touching contact that are not in an island yet
                uint64_t rem = (uint64_t)misc[0] | ((uint64_t)misc[1] << 32);
                int nord = 0;
                while (rem) {
                    const int seed = 63 - __clzll((long long)rem);
                    int sp = 0;
                    stack[sp++] = (uint8_t)seed;
                    rem &= ~(1ull << seed);
                    while (sp > 0) {
                        const int b = stack[--sp];
                        label[b] = (uint8_t)seed;
                        for (int t = head[b]; t != 0xff;) {
                            const int ta = t_a[t], tb = t_b[t];
                            const int nx = (ta == b) ? nxt_a[t] : nxt_b[t];
                            if (!taken[t]) {
                                taken[t] = 1;
                                ord[nord] = (uint16_t)t;
                                const int l = 1 + max((int)lastlvl[ta], (int)lastlvl[tb]);
                                lvl[nord] = (uint8_t)l;
                                lastlvl[ta] = (uint8_t)l;
                                lastlvl[tb] = (uint8_t)l;
                                L = max(L, l);
                                ++nord;
                                const int other = (ta == b) ? tb : ta;
                                if ((rem >> other) & 1) {
                                    rem &= ~(1ull << other);
                                    stack[sp++] = (uint8_t)other;
                                }
                            }
                            t = nx;
                        }
                    }
                }
            }
            nlev = g.shfl(L, 0);
        }
        g.sync();
    }

    // ---- phase 5: contact solver, velocity part -------------------------------------------------
    // Order position k is owned by lane k % G.  The first G positions (all of them, normally)
    // live in that lane's registers for the whole solve; later ones go through shared memory.
    const bool has = g.gl < tc;
    int ka = 0, kb = 0, klv = 0, kt = 0, kisl = 0;
    float knx = 1.0f, kny = 0.0f, knI = 0.0f, ktI = 0.0f;
    if (tc > 0) {
        uint8_t* t_a = S.t_a<NC>(); uint8_t* t_b = S.t_b<NC>();
        float* t_nx = S.t_nx<NC>(); float* t_ny = S.t_ny<NC>();
        float* t_nI = S.t_nI<NC>(); float* t_tI = S.t_tI<NC>();
        uint16_t* ord = S.ord<NC>(); uint8_t* lvl = S.lvl<NC>();
        const float mass_n = P.normal_mass, mass_t = P.normal_mass;
        // b2ContactSolver ctor + InitializeVelocityConstraints: world manifold at the
        // pre-integration positions, impulses scaled by dtRatio
        const float ratio = (es.x == 0) ? 0.0f : P.dt_ratio;  // inv_dt0 == 0 on a world's first step
        for (int k = g.gl; k < tc; k += G) {
            const int t = ord[k];
            const int a = t_a[t], b = t_b[t];
            float nx = 1.0f, ny = 0.0f;
            const float dx = px[b] - px[a], dy = py[b] - py[a];
            // b2DistanceSquared(pointA, pointB) is (A - B).(A - B); squares are sign-blind
            if ((dx * dx + dy * dy) > B2_EPSILON * B2_EPSILON) { nx = dx; ny = dy; b2normalize(nx, ny); }
            float nI = 0.0f, tI = 0.0f;
            if (P.warm_starting) { nI = ratio * t_nI[t]; tI = ratio * t_tI[t]; }
            if (k < G) {
                kt = t; ka = a; kb = b; klv = lvl[k]; kisl = label[a];
                knx = nx; kny = ny; knI = nI; ktI = tI;
            } else {
                t_nx[t] = nx; t_ny[t] = ny; t_nI[t] = nI; t_tI[t] = tI;
            }
        }
        if (nlev == 1) {
            // independent contacts: warm start + all iterations without leaving registers
            if (has) {
                float vax = vx[ka], vay = vy[ka], vbx = vx[kb], vby = vy[kb];
                {
                    const float tx = kny, ty = -knx;
                    const float Px = knI * knx + ktI * tx, Py = knI * kny + ktI * ty;
                    vax -= P.inv_mass * Px; vay -= P.inv_mass * Py;
                    vbx += P.inv_mass * Px; vby += P.inv_mass * Py;
                }
                for (int it = 0; it < P.vel_iters; ++it)
                    solve_velocity(knx, kny, P.friction, mass_n, mass_t, P.inv_mass, knI, ktI, vax, vay, vbx, vby);
                vx[ka] = vax; vy[ka] = vay; vx[kb] = vbx; vy[kb] = vby;
            }
            for (int k = g.gl + G; k < tc; k += G) {
                const int t = ord[k];
                const int a = t_a[t], b = t_b[t];
                const float nx = t_nx[t], ny = t_ny[t];
                float nI = t_nI[t], tI = t_tI[t];
                float vax = vx[a], vay = vy[a], vbx = vx[b], vby = vy[b];
                {
                    const float tx = ny, ty = -nx;
                    const float Px = nI * nx + tI * tx, Py = nI * ny + tI * ty;
                    vax -= P.inv_mass * Px; vay -= P.inv_mass * Py;
                    vbx += P.inv_mass * Px; vby += P.inv_mass * Py;
                }
                for (int it = 0; it < P.vel_iters; ++it)
                    solve_velocity(nx, ny, P.friction, mass_n, mass_t, P.inv_mass, nI, tI, vax, vay, vbx, vby);
                vx[a] = vax; vy[a] = vay; vx[b] = vbx; vy[b] = vby;
                t_nI[t] = nI; t_tI[t] = tI;
            }
            g.sync();
        } else {
            // WarmStart, in order
            for (int lev = 1; lev <= nlev; ++lev) {
                if (has && klv == lev) {
                    const float tx = kny, ty = -knx;
                    const float Px = knI * knx + ktI * tx, Py = knI * kny + ktI * ty;
                    vx[ka] -= P.inv_mass * Px; vy[ka] -= P.inv_mass * Py;
                    vx[kb] += P.inv_mass * Px; vy[kb] += P.inv_mass * Py;
                }
                for (int k = g.gl + G; k < tc; k += G) {
                    if (lvl[k] != lev) continue;
                    const int t = ord[k];
                    const int a = t_a[t], b = t_b[t];
                    const float nx = t_nx[t], ny = t_ny[t], nI = t_nI[t], tI = t_tI[t];
                    const float tx = ny, ty = -nx;
                    const float Px = nI * nx + tI * tx, Py = nI * ny + tI * ty;
                    vx[a] -= P.inv_mass * Px; vy[a] -= P.inv_mass * Py;
                    vx[b] += P.inv_mass * Px; vy[b] += P.inv_mass * Py;
                }
                g.sync();
            }
            for (int it = 0; it < P.vel_iters; ++it) {
                for (int lev = 1; lev <= nlev; ++lev) {
                    if (has && klv == lev) {
                        float vax = vx[ka], vay = vy[ka], vbx = vx[kb], vby = vy[kb];
                        solve_velocity(knx, kny, P.friction, mass_n, mass_t, P.inv_mass, knI, ktI, vax, vay, vbx, vby);
                        vx[ka] = vax; vy[ka] = vay; vx[kb] = vbx; vy[kb] = vby;
                    }
                    for (int k = g.gl + G; k < tc; k += G) {
                        if (lvl[k] != lev) continue;
                        const int t = ord[k];
                        const int a = t_a[t], b = t_b[t];
                        float nI = t_nI[t], tI = t_tI[t];
                        float vax = vx[a], vay = vy[a], vbx = vx[b], vby = vy[b];
                        solve_velocity(t_nx[t], t_ny[t], P.friction, mass_n, mass_t, P.inv_mass, nI, tI, vax, vay,
                                       vbx, vby);
                        vx[a] = vax; vy[a] = vay; vx[b] = vbx; vy[b] = vby;
                        t_nI[t] = nI; t_tI[t] = tI;
                    }
                    g.sync();
                }
            }
        }
        // StoreImpulses -> manifold (next step's warm start)
        uint16_t* t_slot = S.t_slot<NC>();
        if (has) c_imp[t_slot[kt]] = make_float2(knI, ktI);
        for (int k = g.gl + G; k < tc; k += G) { const int t = ord[k]; c_imp[t_slot[t]] = make_float2(t_nI[t], t_tI[t]); }
    }

    // ---- phase 6: integrate positions ------------------------------------------------------------
    float c0x[APL], c0y[APL];
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        c0x[s] = cx[s]; c0y[s] = cy[s];
        float v_x = vx[i], v_y = vy[i];
        const float trx = P.h * v_x, try_ = P.h * v_y;
        if ((trx * trx + try_ * try_) > B2_MAX_TRANSLATION * B2_MAX_TRANSLATION) {
            const float ratio = B2_MAX_TRANSLATION / sqrtf(trx * trx + try_ * try_);
            v_x *= ratio; v_y *= ratio;
        }
        cx[s] += P.h * v_x;
        cy[s] += P.h * v_y;
        wx[s] = v_x; wy[s] = v_y;
    }
    g.sync();
#pragma unroll
    for (int s = 0; s < APL; ++s) { px[g.gl + s * G] = cx[s]; py[g.gl + s * G] = cy[s]; }
    g.sync();

    // ---- phase 7: contact solver, position part (per-island early exit) ---------------------------
    {
        uint8_t* isl_act = S.isl_act<NC>(); uint8_t* isl_bad = S.isl_bad<NC>();
        if (tc > 0) {
            uint8_t* t_a = S.t_a<NC>(); uint8_t* t_b = S.t_b<NC>();
            uint16_t* ord = S.ord<NC>(); uint8_t* lvl = S.lvl<NC>();
            for (int it = 0; it < P.pos_iters; ++it) {
                for (int lev = 1; lev <= nlev; ++lev) {
                    if (has && klv == lev && isl_act[kisl]) {
                        float cax = px[ka], cay = py[ka], cbx = px[kb], cby = py[kb];
                        const float sep = solve_position(P.radius, P.k_sum, P.inv_mass, cax, cay, cbx, cby);
                        px[ka] = cax; py[ka] = cay; px[kb] = cbx; py[kb] = cby;
                        // island not solved while min(0, separations) < -3 * linearSlop
                        if (!(b2min(0.0f, sep) >= -3.0f * B2_LINEAR_SLOP)) isl_bad[kisl] = 1;
                    }
                    for (int k = g.gl + G; k < tc; k += G) {
                        if (lvl[k] != lev) continue;
                        const int t = ord[k];
                        const int a = t_a[t], b = t_b[t];
                        const int isl = label[a];
                        if (!isl_act[isl]) continue;
                        float cax = px[a], cay = py[a], cbx = px[b], cby = py[b];
                        const float sep = solve_position(P.radius, P.k_sum, P.inv_mass, cax, cay, cbx, cby);
                        px[a] = cax; py[a] = cay; px[b] = cbx; py[b] = cby;
                        if (!(b2min(0.0f, sep) >= -3.0f * B2_LINEAR_SLOP)) isl_bad[isl] = 1;
                    }
                    g.sync();
                }
                bool any_bad = false;
#pragma unroll
                for (int s = 0; s < APL; ++s) {
                    const int i = g.gl + s * G;
                    const uint8_t bad = isl_bad[i];
                    isl_act[i] = bad;
                    isl_bad[i] = 0;
                    any_bad |= bad != 0;
                }
                g.sync();
                if (!g.ballot(any_bad)) break;
            }
#pragma unroll
            for (int s = 0; s < APL; ++s) { cx[s] = px[g.gl + s * G]; cy[s] = py[g.gl + s * G]; }
        } else if (P.pos_iters > 0) {
#pragma unroll
            for (int s = 0; s < APL; ++s) isl_act[g.gl + s * G] = 0;
        }
    }

    // ---- phase 8: sleeping (b2Island::Solve tail) --------------------------------------------------
    {
        bool cand = false;
        const float tol2 = B2_LINEAR_SLEEP_TOLERANCE * B2_LINEAR_SLEEP_TOLERANCE;
#pragma unroll
        for (int s = 0; s < APL; ++s) {
            if ((wx[s] * wx[s] + wy[s] * wy[s]) > tol2) slp[s] = 0.0f;
            else slp[s] += P.h;
            cand |= valid[s] && slp[s] >= B2_TIME_TO_SLEEP;
        }
        if (g.ballot(cand)) {
            uint32_t* isl_min = S.isl_min<NC>();
            uint8_t* isl_act = S.isl_act<NC>();
#pragma unroll
            for (int s = 0; s < APL; ++s) isl_min[g.gl + s * G] = 0x7f7fffffu;  // b2_maxFloat
            g.sync();
#pragma unroll
            for (int s = 0; s < APL; ++s)
                if (valid[s]) atomicMin(&isl_min[label[g.gl + s * G]], __float_as_uint(slp[s]));
            g.sync();
#pragma unroll
            for (int s = 0; s < APL; ++s) {
                const int isl = label[g.gl + s * G];
                const bool solved = (P.pos_iters > 0) && !isl_act[isl];
                if (valid[s] && __uint_as_float(isl_min[isl]) >= B2_TIME_TO_SLEEP && solved) {
                    // SetAwake(false); the next ApplyForce(wake=True) wakes the body again
                    slp[s] = 0.0f; wx[s] = 0.0f; wy[s] = 0.0f;
                }
            }
        }
    }

    // ---- phase 9: SynchronizeFixtures -> b2DynamicTree::MoveProxy -----------------------------------
    uint64_t moved = 0;
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        const float r = P.radius;
        const float lox = b2min(c0x[s] - r, cx[s] - r), loy = b2min(c0y[s] - r, cy[s] - r);
        const float hix = b2max(c0x[s] + r, cx[s] + r), hiy = b2max(c0y[s] + r, cy[s] + r);
        const bool contains = fatr[s].x <= lox && fatr[s].y <= loy && hix <= fatr[s].z && hiy <= fatr[s].w;
        const bool mv = valid[s] && !contains;
        if (mv) {
            float nlx = lox - B2_AABB_EXTENSION, nly = loy - B2_AABB_EXTENSION;
            float nhx = hix + B2_AABB_EXTENSION, nhy = hiy + B2_AABB_EXTENSION;
            const float dx = B2_AABB_MULTIPLIER * (cx[s] - c0x[s]), dy = B2_AABB_MULTIPLIER * (cy[s] - c0y[s]);
            if (dx < 0.0f) nlx += dx; else nhx += dx;
            if (dy < 0.0f) nly += dy; else nhy += dy;
            fatr[s] = make_float4(nlx, nly, nhx, nhy);
            flx[i] = nlx; fly[i] = nly; fhx[i] = nhx; fhy[i] = nhy;
        }
        moved |= (uint64_t)g.ballot(mv) << (s * G);
    }
    g.sync();

    // ---- phase 10: FindNewContacts ------------------------------------------------------------------
    if (moved) cnt = find_new_contacts<G, APL>(g, S, P, moved, alive, cnt, c_ab, c_imp, overflow_c);

    // ---- phase 11: rewards (mvmnt.py:160-179), time/done (mvmnt.py:134-136) ---------------------------
    const int step = es.x + 1;
    const bool done = step >= P.done_step;
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        if (!valid[s]) continue;
        const size_t gi = (size_t)env * N + i;
        const bool col = (adj_lo[i] | adj_hi[i]) != 0;
        float rew = -1.0f;
        if (!col) {
            const float2 tg = P.targets[(size_t)env * P.T + P.target_idx[i]];
            const float dx = tg.x - cx[s], dy = tg.y - cy[s];
            const float d2 = dx * dx + dy * dy;
            if (P.reward_mode == MACM_REWARD_LINEAR) rew = (-sqrtf(d2) / 35.0f) + 1.0f;
            else rew = (d2 < P.binary_thr) ? 1.0f : 0.0f;
        }
        P.rewards[gi] = rew;
        P.collided[gi] = (uint8_t)col;
        // ---- phase 12: write state back ----
        P.posvel[gi] = make_float4(cx[s], cy[s], wx[s], wy[s]);
        P.angsleep[gi] = make_float2(ang[s], slp[s]);
        P.fat[gi] = fatr[s];
    }
    {
        const bool oc = g.ballot(overflow_c) != 0, ot = g.ballot(overflow_t) != 0;
        if (g.gl == 0) {
            const int flags = (es.y & ~MACM_ENV_FRESH) | (oc ? MACM_ENV_CONTACT_OVERFLOW : 0) |
                              (ot ? MACM_ENV_TOUCH_OVERFLOW : 0);
            P.c_cnt[env] = cnt;
            P.done[env] = (uint8_t)done;
            P.env_state[env] = make_int4(step, flags, tc, es.w);
        }
    }

    // ---- phase 13: observations (mvmnt.py:181-222) ------------------------------------------------
    flock_observe<G, APL>(g, S, P, env, ang);
}

// get_obs() alone
template <int G, int APL>
__global__ void __launch_bounds__(128) macm_flock_observe_kernel(const __grid_constant__ SimConst P)
{
    constexpr int NC = G * APL;
    constexpr int GPW = 32 / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Grp<G> g;
    const int slot_in_block = (threadIdx.x >> 5) * GPW + (threadIdx.x & 31) / G;
    const int env = blockIdx.x * (blockDim.x / 32) * GPW + slot_in_block;
    if (env >= P.E) return;
    EnvS S;
    S.TC = P.TC;
    S.base = smem_raw + (size_t)slot_in_block * Lay<NC>::bytes(P.TC);
    float ang[APL];
#pragma unroll
    for (int s = 0; s < APL; ++s) {
        const int i = g.gl + s * G;
        const bool v = i < P.N;
        const size_t gi = (size_t)env * P.N + (v ? i : 0);
        const float4 pv = P.posvel[gi];
        ang[s] = P.angsleep[gi].x;
        S.px<NC>()[i] = v ? pv.x : 3.0e30f;
        S.py<NC>()[i] = v ? pv.y : 3.0e30f;
    }
    g.sync();
    flock_observe<G, APL>(g, S, P, env, ang);
}

// Body creation for every agent (mvmnt.py:61-76): fat AABB = tight +- b2_aabbExtension, awake,
// sleep time 0, no contacts, first-step flag.  One thread per agent.
__global__ void macm_reset_kernel(const __grid_constant__ SimConst P)
{
    const size_t gi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)P.E * P.N;
    if (gi >= total) return;
    const float4 pv = P.posvel[gi];
    const float r = P.radius;
    P.fat[gi] = make_float4((pv.x - r) - B2_AABB_EXTENSION, (pv.y - r) - B2_AABB_EXTENSION,
                            (pv.x + r) + B2_AABB_EXTENSION, (pv.y + r) + B2_AABB_EXTENSION);
    float2 as = P.angsleep[gi];
    as.y = 0.0f;
    P.angsleep[gi] = as;
    P.rewards[gi] = 0.0f;
    P.collided[gi] = 0;
    if (P.kind == MACM_ENV_TDM) P.tdm[gi] = make_float4(P.init_health, __int_as_float(0), __int_as_float(0), 1.0f);
    if (gi % P.N == 0) {
        const size_t e = gi / P.N;
        P.c_cnt[e] = 0;
        P.env_state[e] = make_int4(0, MACM_ENV_FRESH, 0, -1);
        P.done[e] = 0;
    }
}

template <int G, int APL>
cudaError_t launch_flock(const SimConst& P, const LaunchCfg& cfg, const void* actions, cudaStream_t s, bool observe_only)
{
    if (observe_only) macm_flock_observe_kernel<G, APL><<<cfg.blocks, cfg.threads, cfg.smem_bytes, s>>>(P);
    else macm_flock_step_kernel<G, APL><<<cfg.blocks, cfg.threads, cfg.smem_bytes, s>>>(P, actions);
    return cudaGetLastError();
}

template <int G, int APL>
cudaError_t prepare_flock(const LaunchCfg& cfg, int* blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(macm_flock_step_kernel<G, APL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         cfg.smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(macm_flock_observe_kernel<G, APL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             cfg.smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, macm_flock_step_kernel<G, APL>, cfg.threads,
                                                         cfg.smem_bytes);
}

}  // namespace

// lanes per env / agents per lane for a given N
static void pick_shape(int N, int* G, int* APL)
{
    if (N <= 4) { *G = 4; *APL = 1; }
    else if (N <= 8) { *G = 8; *APL = 1; }
    else if (N <= 16) { *G = 16; *APL = 1; }
    else if (N <= 32) { *G = 32; *APL = 1; }
    else { *G = 32; *APL = 2; }
}

cudaError_t macm_launch_cfg(const SimConst& P, LaunchCfg* cfg)
{
    pick_shape(P.N, &cfg->G, &cfg->APL);
    cfg->threads = 128;
    cfg->envs_per_block = (cfg->threads / 32) * (32 / cfg->G);
    cfg->blocks = (P.E + cfg->envs_per_block - 1) / cfg->envs_per_block;
    const int NC = cfg->G * cfg->APL;
    int per_env = 0;
    switch (NC) {
        case 4: per_env = Lay<4>::bytes(P.TC); break;
        case 8: per_env = Lay<8>::bytes(P.TC); break;
        case 16: per_env = Lay<16>::bytes(P.TC); break;
        case 32: per_env = Lay<32>::bytes(P.TC); break;
        default: per_env = Lay<64>::bytes(P.TC); break;
    }
    cfg->smem_bytes = per_env * cfg->envs_per_block;
    return cfg->smem_bytes <= 227 * 1024 ? cudaSuccess : cudaErrorInvalidConfiguration;
}

#define DISPATCH_SHAPE(CALL)                                   \
    switch (cfg.G * 8 + cfg.APL) {                             \
        case 4 * 8 + 1: return CALL(4, 1);                     \
        case 8 * 8 + 1: return CALL(8, 1);                     \
        case 16 * 8 + 1: return CALL(16, 1);                   \
        case 32 * 8 + 1: return CALL(32, 1);                   \
        case 32 * 8 + 2: return CALL(32, 2);                   \
        default: return cudaErrorInvalidConfiguration;         \
    }

cudaError_t macm_prepare_kernels(const SimConst& P, const LaunchCfg& cfg, int* blocks_per_sm)
{
    (void)P;
#define CALL(G_, A_) prepare_flock<G_, A_>(cfg, blocks_per_sm)
    DISPATCH_SHAPE(CALL)
#undef CALL
}

cudaError_t macm_launch_step(const SimConst& P, const LaunchCfg& cfg, const void* actions, cudaStream_t s)
{
#define CALL(G_, A_) launch_flock<G_, A_>(P, cfg, actions, s, false)
    DISPATCH_SHAPE(CALL)
#undef CALL
}

cudaError_t macm_launch_observe(const SimConst& P, const LaunchCfg& cfg, cudaStream_t s)
{
#define CALL(G_, A_) launch_flock<G_, A_>(P, cfg, nullptr, s, true)
    DISPATCH_SHAPE(CALL)
#undef CALL
}

cudaError_t macm_launch_reset(const SimConst& P, cudaStream_t s)
{
    const size_t total = (size_t)P.E * P.N;
    const int threads = 256;
    macm_reset_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(P);
    return cudaGetLastError();
}
