// glue.cu -- node glue on the way into the solve: the continuous-yaw accumulator of BLUEROV2_DOB::solve
// (bluerov2_dobmpc/src/bluerov2_dob.cpp:272-304; identical in bluerov2_ampc.cpp:285-317).
//
// The pose callback hands the node a yaw in (-pi, pi] (tf getRPY); the OCP needs a continuous angle, so the node keeps
// `pre_yaw` and `yaw_sum` and adds the shortest signed difference every tick.  Both are declared FLOAT in the reference
// (bluerov2_dob.h:234-236), so x0[psi] carries float32 rounding of the accumulated yaw -- reproduced here, conversions
// included (double - float -> float, float + float, psi -> float), because x0 is an input of the solve and parity is
// judged on u.  One thread per instance; state = (pre_yaw, yaw_sum) as float2.
#include "engine.h"

namespace br2 {

__global__ void __launch_bounds__(128) yaw_unwrap_kernel(int B, float2* state, double* x0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double TWO_PI = 2 * 3.14159265358979323846;     // 2*M_PI
    float2 st = state[i];
    const float pre_yaw = st.x;
    const double psi = x0[(size_t)i * NX + 5];
    float yaw_diff;
    if (pre_yaw >= 0 && psi >= 0) {
        yaw_diff = (float)(psi - pre_yaw);
    } else if (pre_yaw >= 0 && psi < 0) {
        if (TWO_PI + psi - pre_yaw >= pre_yaw + fabs(psi)) yaw_diff = (float)(-(pre_yaw + fabs(psi)));
        else yaw_diff = (float)(TWO_PI + psi - pre_yaw);
    } else if (pre_yaw < 0 && psi >= 0) {
        if (TWO_PI - psi + pre_yaw >= fabsf(pre_yaw) + psi) yaw_diff = (float)(fabsf(pre_yaw) + psi);
        else yaw_diff = (float)(-(TWO_PI - psi + pre_yaw));
    } else {
        yaw_diff = (float)(psi - pre_yaw);
    }
    st.y = st.y + yaw_diff;            // yaw_sum: float + float
    st.x = (float)psi;                 // pre_yaw = local_euler.psi
    state[i] = st;
    x0[(size_t)i * NX + 5] = (double)st.y;     // acados_in.x0[psi] = yaw_sum (:312)
}

void launch_yaw_unwrap(int B, float* state, double* x0, cudaStream_t s)
{
    yaw_unwrap_kernel<<<(B + 127) / 128, 128, 0, s>>>(B, reinterpret_cast<float2*>(state), x0);
}

}  // namespace br2
