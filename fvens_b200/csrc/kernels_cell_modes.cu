/* Instantiations of the gradient + limiter pass with features switched on (cell_kernel.cuh): multi-GPU pushes and waits
 * (CM_DIST), caller-ordered state (CM_PERM), both, and - for the headline numerics only, kept as a measurable experiment -
 * the looping form (CM_LOOP). A translation unit of its own so that it compiles in parallel with kernels_cell.cu. */
#include "cell_kernel.cuh"

namespace fvg {

int launch_cell_kernel_modes(int grad, int lim, int mode, const CellArgs &b, cudaStream_t s)
{
	const int nt = b.tile1 - b.tile0;
#define C(G,L) if(grad == G && lim == L) { \
		if(mode == CM_DIST) return launch_cell_grid<G,L,false,CM_DIST>(b, nt, s); \
		if(mode == CM_PERM) return launch_cell_grid<G,L,false,CM_PERM>(b, nt, s); \
		if(mode == (CM_DIST | CM_PERM)) return launch_cell_grid<G,L,false,CM_DIST | CM_PERM>(b, nt, s); }
	C(GM_ZERO,LM_NONE) C(GM_ZERO,LM_BJ) C(GM_ZERO,LM_VENKAT)
	C(GM_GG,LM_NONE) C(GM_GG,LM_BJ) C(GM_GG,LM_VENKAT)
	C(GM_WLS,LM_NONE) C(GM_WLS,LM_BJ) C(GM_WLS,LM_VENKAT)
#undef C
	if((mode & CM_LOOP) && grad == GM_WLS && lim == LM_VENKAT && !b.src_idx && !b.ucopy)
		return launch_cell_grid<GM_WLS,LM_VENKAT,false,CM_LOOP | CM_DIST>(b, 0, s);
	if(mode & CM_LOOP) return launch_cell_kernel_modes(grad, lim, mode & ~CM_LOOP ? mode & ~CM_LOOP : CM_DIST, b, s);
	set_error("cell kernel: unsupported gradient/limiter combination");
	return FVG_ERR_INVALID;
}

} // namespace fvg
