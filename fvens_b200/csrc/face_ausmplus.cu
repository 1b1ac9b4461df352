/* Face-kernel instantiations for one inviscid flux (see face_kernel.cuh). One translation unit per
 * flux so that the nine (reconstruction x viscosity) variants of each flux compile in parallel. */
#include "face_kernel.cuh"
namespace fvg {
int launch_face_ausmplus(int recon, int visc, const FaceArgs &a, cudaStream_t s) { return launch_flux<FLUX_AUSMPLUS>(recon, visc, a, s); }
}
