// Coefficient pass, cell half 1: eval_crust_eos per (zone, knot) -> lnCp, lnGamma tables and the rho_bar(1) fix-up
// (NScool_create_model.f:344-369, :386-397).
// alignment barriers between the components only, not inside ee_exchange (4.04 -> 4.00 ms).
#define DS_NO_PHASE_INNER
#define DS_MICRO_PART DS_PART_EOS
#define DS_MICRO_FN ds_launch_micro_eos
#ifdef DS_MICRO_SCAN
#define DS_MICRO_GS(X) X(3) X(4) X(5) X(6) X(8) X(7)
#else
#define DS_MICRO_GS(X) X(7)
#endif
#define DS_MICRO_FAST 1
#include "micro_part.cuh"
