// Coefficient pass, cells, fused form (EOS + neutrino emissivity in one kernel): the round-1 kernel, kept selectable
// (DSTAR_MICRO_SPLIT=0) as the comparison for the split pass.
#define DS_MICRO_PART DS_PART_CELL
#define DS_MICRO_FN ds_launch_micro_cell
#define DS_MICRO_GS(X) X(6) X(7)
#define DS_MICRO_FAST 1
#include "micro_part.cuh"
