// macm_aux.cu -- kernels beside the step: the on-device initial-state sampler
// (Flock.__init__ / TDM.__init__ distributions, mvmnt.py:48-52,62-64; combat.py:84-86) and the
// scripted actors of test_scripts/bots.py (the `actions=None` mode of mvmnt.py:86-92).
#include "macm_sim.h"

namespace {

__global__ void macm_sample_kernel(const __grid_constant__ SimConst P, uint64_t seed, double spread, double sx,
                                   double sy, double tmin, double tmax, double width, double height)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t EN = (uint64_t)P.E * P.N, ET = (uint64_t)P.E * P.T;
    if (idx < EN) {
        const uint64_t gidx = idx + (uint64_t)P.env_base * P.N;   // global agent index
        const Philox r(seed, gidx, 0u, 0u);
        const Philox r2(seed, gidx, 0u, 1u);
        const int i = (int)(idx % P.N);
        double x, y;
        if (P.kind == MACM_ENV_FLOCK) {
            x = spread * (r.u53(0) - 0.5) + sx;   // mvmnt.py:62-63
            y = spread * (r.u53(1) - 0.5) + sy;
        } else {
            x = r.u53(0) * ((double)P.team[i] + width / 2);  // combat.py:84-85
            y = r.u53(1) * height;
        }
        const double a = (-1.0 + 2.0 * r2.u53(0)) * NP_PI;   // random.uniform(-1, 1) * np.pi
        P.posvel[idx] = make_float4((float)x, (float)y, 0.0f, 0.0f);
        P.angsleep[idx] = make_float2((float)a, 0.0f);
    } else if (idx < EN + ET) {
        const uint64_t t = idx - EN;
        const Philox r(seed, t + (uint64_t)P.env_base * P.T, 1u, 0u);
        const double ang = 2 * NP_PI * r.u53(0);              // mvmnt.py:50-52
        const double dist = tmin + r.u53(1) * (tmax - tmin);
        reinterpret_cast<float2*>(const_cast<float2*>(P.targets))[t] =
            make_float2((float)(dist * cos(ang)), (float)(dist * sin(ang)));
    }
}

// bots.py: idle/forward/rotate/diag (bots:19-29), flock (bots:37-61), plus U{0,1,2}^3 x U{0,1}
__global__ void macm_bot_kernel(const __grid_constant__ SimConst P, int policy, uint64_t seed, uint32_t* out)
{
    const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (uint64_t)P.E * P.N) return;
    uint32_t a0 = 1, a1 = 1, a2 = 1, a3 = 0;
    switch (policy) {
        case MACM_BOT_FORWARD: a0 = 2; break;
        case MACM_BOT_ROTATE: a2 = 2; break;
        case MACM_BOT_DIAG: a0 = 2; a1 = 2; break;
        case MACM_BOT_RANDOM: {
            const int step = P.env_state[gi / P.N].x;
            const Philox r(seed, gi + (uint64_t)P.env_base * P.N, 2u, (uint32_t)step);
            a0 = (uint32_t)(((uint64_t)r.c[0] * 3u) >> 32);
            a1 = (uint32_t)(((uint64_t)r.c[1] * 3u) >> 32);
            a2 = (uint32_t)(((uint64_t)r.c[2] * 3u) >> 32);
            a3 = P.kind == MACM_ENV_TDM ? (r.c[3] >> 31) : 0u;
            break;
        }
        case MACM_BOT_FLOCK: {
            // steer towards the target node: position = [r, theta] or [r, cos, sin]
            float r, th_sign, ahead;
            if (P.coord == MACM_COORD_POLAR) {
                const float4 o = reinterpret_cast<const float4*>(P.obs)[gi];
                r = o.z;
                th_sign = (o.w > 0.0f) ? 1.0f : (o.w < 0.0f ? -1.0f : 0.0f);          // np.sign(theta)
                ahead = fabsf(o.w) < (float)(NP_PI / 4) ? 1.0f : 0.0f;                 // |theta| < pi/4
            } else {
                const float* o = P.obs + gi * 6;
                r = o[3];
                th_sign = (o[5] > 0.0f) ? 1.0f : (o[5] < 0.0f ? -1.0f : 0.0f);        // np.sign(sin)
                ahead = o[4] > 0.70710678f ? 1.0f : 0.0f;                              // cos > cos(pi/4)
            }
            if (!(r < 1.0f)) { a2 = (uint32_t)(th_sign + 1.0f); a0 = (uint32_t)(ahead + 1.0f); }
            break;
        }
        default: break;
    }
    out[gi] = a0 | (a1 << 8) | (a2 << 16) | (a3 << 24);
}

}  // namespace

cudaError_t macm_launch_sample(const SimConst& P, uint64_t seed, double start_spread, double start_x, double start_y,
                               double tmin, double tmax, double width, double height, cudaStream_t s)
{
    const uint64_t total = (uint64_t)P.E * P.N + (uint64_t)P.E * P.T;
    const int threads = 256;
    macm_sample_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(P, seed, start_spread, start_x,
                                                                                       start_y, tmin, tmax, width, height);
    return cudaGetLastError();
}

cudaError_t macm_launch_bot(const SimConst& P, int policy, uint64_t seed, void* actions_out, cudaStream_t s)
{
    const uint64_t total = (uint64_t)P.E * P.N;
    const int threads = 256;
    macm_bot_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(P, policy, seed,
                                                                                    (uint32_t*)actions_out);
    return cudaGetLastError();
}
