// k_motion.cu -- stage 1: odometry propagation for every particle.
// Reference: Robot.imu_update robot.py:45-57 with the loader callbacks
// IntelIMUData.py:22-36 (absolute), IntelRawIMUData.py:33-55 (velocity; same
// shape in Aces/Freid*/Obero/Bele), DefaultIMUData.py:25-54 (unicycle).
// One thread per particle; 12 doubles in, 12 out.
#include "common.cuh"

struct MotionArgs {
    int family;
    double u[4], par[4], dt;
};

__device__ __forceinline__ void mat3mul(const double *a, const double *b, double *o)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            o[3 * r + c] = (a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c]) + a[3 * r + 2] * b[6 + c];
}

__global__ void __launch_bounds__(128) motion_kernel(RbCtx c, MotionArgs a)
{
    const double PI = 3.14159265358979323846;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.N) return;
    double *pose = c.pose + 3 * (size_t)i, *cov = c.cov + 9 * (size_t)i;
    double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Q[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, np_[3];
    double x = pose[0], y = pose[1], th = pose[2], dt = a.dt;
    if (a.family == 0) {                      // IntelIMUData.py:22-36 (callbacks' contents swapped, kept)
        np_[0] = a.u[0]; np_[1] = a.u[1]; np_[2] = a.u[2];
        F[0] = 0.01 * 0.01; F[4] = 0.01 * 0.01; F[8] = (0.2 * PI / 180) * (0.2 * PI / 180);
        Q[0] = Q[4] = Q[8] = 1.0;
        Q[2] = a.u[0] - x;
        Q[5] = a.u[1] - y;
    } else if (a.family == 1) {               // IntelRawIMUData.py:33-55
        np_[0] = x + a.u[0] * dt; np_[1] = y + a.u[1] * dt; np_[2] = th + a.u[2] * dt;
        double q0 = a.par[0] + a.par[1] * fabs(a.u[0]) * dt, q1 = a.par[0] + a.par[1] * fabs(a.u[1]) * dt,
               q2 = a.par[2] + a.par[3] * fabs(a.u[2]) * dt;
        Q[0] = fabs(q0 * q0); Q[4] = fabs(q1 * q1); Q[8] = fabs(q2 * q2);
    } else {                                  // DefaultIMUData.py:25-54
        double t2 = th + dt * a.u[1];
        double ct, st, cp, sp;
        rb_sincos(t2, &st, &ct);
        rb_sincos(th, &sp, &cp);
        np_[0] = x + dt * a.u[0] * ct;
        np_[1] = y + dt * a.u[0] * st;
        np_[2] = t2;
        F[2] = dt * a.u[0] * cp;
        F[5] = dt * a.u[0] * sp;
        double g0 = dt * cp, g1 = dt * sp, g2 = dt, m0 = 0.05 * 0.05, m1 = (PI / 180 / 2) * (PI / 180 / 2);
        Q[0] = fabs((g0 * m0) * g0) + 0.01 * 0.01;
        Q[1] = fabs((g0 * m0) * g1);
        Q[3] = fabs((g1 * m0) * g0);
        Q[4] = fabs((g1 * m0) * g1) + 0.01 * 0.01;
        Q[8] = fabs((g2 * m1) * g2) + (0.2 * PI / 180) * (0.2 * PI / 180);
    }
    double cv[9], t1[9], Ft[9], t2m[9];
#pragma unroll
    for (int k = 0; k < 9; k++) cv[k] = cov[k];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) Ft[3 * r + q] = F[3 * q + r];
    mat3mul(F, cv, t1);                       // robot.py:50
    mat3mul(t1, Ft, t2m);
#pragma unroll
    for (int k = 0; k < 9; k++) cov[k] = t2m[k] + Q[k];   // robot.py:51
    pose[0] = np_[0]; pose[1] = np_[1]; pose[2] = np_[2];
}

void rb_launch_motion(const RbCtx &c, int family, const double *u, double dt, const double *par, cudaStream_t s)
{
    MotionArgs a;
    a.family = family;
    a.dt = dt;
    for (int k = 0; k < 4; k++) { a.u[k] = u[k]; a.par[k] = par ? par[k] : 0.0; }
    motion_kernel<<<(c.N + 127) / 128, 128, 0, s>>>(c, a);
}
