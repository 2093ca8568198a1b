// plant.cu -- batched nominal plant for device-resident closed-loop studies.
//
// One classical RK4 step of the OCP model (bluerov2_dobmpc/scripts/bluerov2.py:103-137, the same csrc/model.cuh the
// lineariser integrates) per instance, with an optional true disturbance (X, Y, Z, N) added to p[0..3] -- the "wave"
// wrench of applyBodyWrench mode 0 (bluerov2_dob.cpp:774-797) is produced on the fly from per-instance amplitudes and
// phases; mode 1 (:813-816) is the constant `dist`; mode 2 (:818-874) replays an uploaded series row by row.  Also emits what the node's pose callback derives for the EKF (bluerov2_dob.cpp:148-153): the body
// acceleration as the finite difference of the body velocities.  One thread per instance: 12 states, ~0.4 kflop.
#include "engine.h"

namespace br2 {

__global__ void __launch_bounds__(128) plant_kernel(PlantArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B) return;
    double x[NX], u[NU], p[NP];
#pragma unroll
    for (int j = 0; j < NX; j++) x[j] = a.x[(size_t)i * NX + j];
#pragma unroll
    for (int j = 0; j < NU; j++) u[j] = a.u[(size_t)i * NU + j];
#pragma unroll
    for (int j = 0; j < NP; j++) p[j] = a.p[(size_t)i * NP + j];
    if (a.wave_amp) {
        // F = sin(tau) * A,  tau = tau0 + 0.125 * tick  (bluerov2_dob.cpp:774-797: tau advances 0.05 * 2.5 per tick)
        const int tick = a.tick >= 0 ? a.tick : *a.tick_ctr - 1;
        const double sn = sin(a.wave_tau0[i] + 0.125 * tick);
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] += sn * a.wave_amp[(size_t)i * 4 + j];
    }
    if (a.table) {
        // mode 2 (bluerov2_dob.cpp:818-874): the wrench of this tick is a row of the series read from config/force{x,y,z}.txt and
        // torquez.txt, all four at the same counter (fx_counter++); past the end of the series the last row holds (the reference
        // reads beyond its vectors there)
        const int tick = a.tick >= 0 ? a.tick : *a.tick_ctr - 1;
        int row = tick + (a.table_phase ? a.table_phase[i] : 0);
        row = row < 0 ? 0 : (row < a.table_rows ? row : a.table_rows - 1);
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] += a.table[(size_t)row * 4 + j];
    }
    if (a.dist) {
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] += a.dist[(size_t)i * 4 + j];
    }
    ModelConst mc;
    mc.set(p);
    double k[NX], xs[NX], acc[NX];
    Trig t;
    const double h = a.h;
    // k1
    trig_of(x, t); ode(x, u, mc, t, k);
#pragma unroll
    for (int j = 0; j < NX; j++) { acc[j] = k[j] * (1.0 / 6.0); xs[j] = x[j] + 0.5 * h * k[j]; }
    // k2
    trig_of(xs, t); ode(xs, u, mc, t, k);
#pragma unroll
    for (int j = 0; j < NX; j++) { acc[j] += k[j] * (1.0 / 3.0); xs[j] = x[j] + 0.5 * h * k[j]; }
    // k3
    trig_of(xs, t); ode(xs, u, mc, t, k);
#pragma unroll
    for (int j = 0; j < NX; j++) { acc[j] += k[j] * (1.0 / 3.0); xs[j] = x[j] + h * k[j]; }
    // k4
    trig_of(xs, t); ode(xs, u, mc, t, k);
#pragma unroll
    for (int j = 0; j < NX; j++) {
        const double xn = x[j] + h * (acc[j] + k[j] * (1.0 / 6.0));
        a.x[(size_t)i * NX + j] = xn;
        if (a.body_acc && j >= 6) a.body_acc[(size_t)i * 6 + (j - 6)] = (xn - x[j]) / h;
    }
    if (a.lines) a.lines[i] += 1;      // line_number++ (bluerov2_dob.cpp:367)
}

void launch_plant(const PlantArgs& a, cudaStream_t s)
{
    plant_kernel<<<(a.B + 127) / 128, 128, 0, s>>>(a);
}

}  // namespace br2
