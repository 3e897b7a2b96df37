// Coefficient pass, faces: EOS (Gamma, Theta, mu_e) + thermal conductivity per (face, knot) -> lnK table
// (NScool_create_model.f:406-441).
// The face evaluation executes ~3 k instructions per lane (~50 KB of SASS): it stays in the instruction caches without the
// alignment barriers of the evaluators (faces 1.34 -> 1.26 ms without them; measured per part, profiles/r02_notes.md).
#define DS_NO_PHASE_SYNC
#define DS_MICRO_PART DS_PART_FACE
#define DS_MICRO_FN ds_launch_micro_face
#ifdef DS_MICRO_SCAN
#define DS_MICRO_GS(X) X(3) X(4) X(5) X(6) X(7)
#else
#define DS_MICRO_GS(X) X(7)
#endif
#define DS_MICRO_FAST 0
// block schedule: 28 warps per CTA (72 registers): faces 1.14 ms against 1.18 at 24, accretion faces 11.0 against 11.5
#define DS_BLK_W_X 28
#define DS_MICRO_BLOCK_FN ds_launch_micro_face_blocks
#include "micro_part.cuh"
