// pf_sim_f32.cu -- f32 event-loop instantiations of the simulate+weight kernel (DPOMP_SIM_F32, the fast default).
#include "pf_sim.cuh"
namespace dpomp {
cudaError_t launch_sim_f32(const ModelHost& m, int items, const SimLaunch& a, cudaStream_t stream) {
    return launch_sim_typed<float>(m, items, a, stream);
}
}  // namespace dpomp
