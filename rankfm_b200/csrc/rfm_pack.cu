// rfm_pack.cu -- conversion between the reference's array layout (rankfm/rankfm.py:214-244: w_i [I], v_u [U,F],
// v_i [I,F], x_uf [U,P], x_if [I,Q], ...) and the fat-row HBM tables described in rfm_common.cuh.
#include "rfm_kernels.h"

namespace rfm {

// v_u / x_uf here are the staged rows of the OWNED users [T.u0, T.u0+T.Un) (row 0 = user T.u0)
__global__ void pack_users_kernel(const Tables T, const float* __restrict__ v_u, const float* __restrict__ x_uf)
{
    const long long n = (long long)T.Un * T.ldu;
    float* base = T.UT + (long long)T.u0 * T.ldu;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long u = e / T.ldu;
        const int c = (int)(e % T.ldu);
        float v = 0.f;
        if (c < T.F) v = v_u[u * T.F + c];
        else if (c >= T.Fp && c - T.Fp < T.P && T.Pp > 0) v = x_uf[u * T.P + (c - T.Fp)];
        base[e] = v;
    }
}

__global__ void pack_items_kernel(const Tables T, const float* __restrict__ v_i, const float* __restrict__ w_i, const float* __restrict__ x_if)
{
    const long long n = (long long)T.I * T.ldi;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / T.ldi;
        const int c = (int)(e % T.ldi);
        float v = 0.f;
        if (c < T.F) v = v_i[i * T.F + c];
        else if (c == T.Fp) v = w_i[i];
        else if (c >= T.Fp + 4 && c - T.Fp - 4 < T.Q && T.Qp > 0) v = x_if[i * T.Q + (c - T.Fp - 4)];
        T.IT[e] = v;
    }
}

__global__ void unpack_users_kernel(const Tables T, float* __restrict__ v_u)
{
    const long long n = (long long)T.Un * T.F;
    const float* base = T.UT + (long long)T.u0 * T.ldu;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        v_u[e] = base[(e / T.F) * T.ldu + (e % T.F)];
}

__global__ void unpack_items_kernel(const Tables T, float* __restrict__ v_i, float* __restrict__ w_i)
{
    const long long n = (long long)T.I * T.F;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        v_i[e] = T.IT[(e / T.F) * T.ldi + (e % T.F)];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < T.I; i += (long long)gridDim.x * blockDim.x)
        w_i[i] = T.IT[i * T.ldi + T.Fp];
}

// GP <-> (w_if [Q], v_uf [P,F], v_if [Q,F]); inactive blocks are left untouched on the way back
__global__ void pack_globals_kernel(const Tables T, const float* __restrict__ w_if, const float* __restrict__ v_uf, const float* __restrict__ v_if, int n_total)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_total; e += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (e < T.gp_vuf) { if (e < T.Q) v = w_if[e]; }          // inactive blocks are packed too: epoch-end statistics cover all six arrays
        else if (e < T.gp_vif) { const int o = e - T.gp_vuf, p = o / T.Fp, f = o % T.Fp; if (f < T.F) v = v_uf[p * T.F + f]; }
        else { const int o = e - T.gp_vif, q = o / T.Fp, f = o % T.Fp; if (f < T.F) v = v_if[q * T.F + f]; }
        T.GP[e] = v;
    }
}

__global__ void unpack_globals_kernel(const Tables T, float* __restrict__ w_if, float* __restrict__ v_uf, float* __restrict__ v_if)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (T.x_if_any) for (int q = tid; q < T.Q; q += nth) w_if[q] = T.GP[q];
    for (int e = tid; e < T.P * T.F; e += nth) v_uf[e] = T.GP[T.gp_vuf + (e / T.F) * T.Fp + (e % T.F)];
    for (int e = tid; e < T.Q * T.F; e += nth) v_if[e] = T.GP[T.gp_vif + (e / T.F) * T.Fp + (e % T.F)];
}

static inline int grid_for(long long n) { return (int)(n / 256 + 1 > 148 * 8 ? 148 * 8 : n / 256 + 1); }

cudaError_t launch_pack_users(const Tables& T, const float* v_u, const float* x_uf, cudaStream_t st)
{
    pack_users_kernel<<<grid_for((long long)T.Un * T.ldu), 256, 0, st>>>(T, v_u, x_uf);
    return cudaGetLastError();
}
cudaError_t launch_pack_items(const Tables& T, const float* v_i, const float* w_i, const float* x_if, cudaStream_t st)
{
    pack_items_kernel<<<grid_for((long long)T.I * T.ldi), 256, 0, st>>>(T, v_i, w_i, x_if);
    return cudaGetLastError();
}
cudaError_t launch_unpack_users(const Tables& T, float* v_u, cudaStream_t st)
{
    unpack_users_kernel<<<grid_for((long long)T.Un * T.F), 256, 0, st>>>(T, v_u);
    return cudaGetLastError();
}
cudaError_t launch_unpack_items(const Tables& T, float* v_i, float* w_i, cudaStream_t st)
{
    unpack_items_kernel<<<grid_for((long long)T.I * T.F), 256, 0, st>>>(T, v_i, w_i);
    return cudaGetLastError();
}
cudaError_t launch_pack_globals(const Tables& T, const float* w_if, const float* v_uf, const float* v_if, int n_total, cudaStream_t st)
{
    pack_globals_kernel<<<grid_for(n_total), 256, 0, st>>>(T, w_if, v_uf, v_if, n_total);
    return cudaGetLastError();
}
cudaError_t launch_unpack_globals(const Tables& T, float* w_if, float* v_uf, float* v_if, cudaStream_t st)
{
    unpack_globals_kernel<<<8, 256, 0, st>>>(T, w_if, v_uf, v_if);
    return cudaGetLastError();
}

}  // namespace rfm
