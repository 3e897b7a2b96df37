// Coefficient pass, cell half 2: crust neutrino emissivity per (zone, knot) -> lnEnu table, with the shell-Urca term
// (NScool_create_model.f:371-384, :398-403).
// ~2.7 k instructions per lane: no alignment barriers needed (0.97 -> 0.93 ms without them).
#define DS_NO_PHASE_SYNC
#define DS_MICRO_PART DS_PART_NU
#define DS_MICRO_FN ds_launch_micro_nu
#ifdef DS_MICRO_SCAN
#define DS_MICRO_GS(X) X(4) X(5) X(6) X(7) X(8)
#else
#define DS_MICRO_GS(X) X(8)
#endif
#define DS_MICRO_FAST 1
// block schedule: 32 warps per CTA (64 registers, no spill frame): 0.918 ms against 0.929 at 24
#define DS_BLK_W_X 32
#define DS_MICRO_BLOCK_FN ds_launch_micro_nu_blocks
#include "micro_part.cuh"
