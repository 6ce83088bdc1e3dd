// Stitches the per-(ND, NP) kernel tables (inst.cu) into one lookup.  The pair list comes
// from the Makefile through the generated build/pairs.h:  #define FLOU_PAIR_LIST X(1,2) ...
#include "launch.h"
#include "pairs.h"

namespace flou {

#define X(nd, np)                                                                 \
    const StageLauncher *stage_table_##nd##_##np##_p0(int eq, int vol, int cart); \
    const StageLauncher *stage_table_##nd##_##np##_p1(int eq, int vol, int cart); \
    const StageLauncher *stage_table_##nd##_##np##_p2(int eq, int vol, int cart); \
    const EmitLauncher *emit_table_##nd##_##np##_p0(int nv);
FLOU_PAIR_LIST
#undef X

const StageLauncher *get_stage_launcher(int nd, int np, int eq, int vol, int cart)
{
    if (eq < 0 || eq > 1 || vol < 0 || vol > 6) return nullptr;
    // translation unit of the instance (inst.cu): 0 = strong / split, 1 = hybrid (vol 3, 6), 2 = split on Gauss nodes (4, 5)
    const int part = vol < 3 ? 0 : ((vol == 3 || vol == 6) ? 1 : 2);
#define X(a, b)                                                                              \
    if (nd == a && np == b)                                                                  \
        return part == 0 ? stage_table_##a##_##b##_p0(eq, vol, cart)                         \
                         : (part == 1 ? stage_table_##a##_##b##_p1(eq, vol, cart) : stage_table_##a##_##b##_p2(eq, vol, cart));
    FLOU_PAIR_LIST
#undef X
    return nullptr;
}

const EmitLauncher *get_emit_launcher(int nd, int np, int nv)
{
#define X(a, b) if (nd == a && np == b) return emit_table_##a##_##b##_p0(nv);
    FLOU_PAIR_LIST
#undef X
    return nullptr;
}

}  // namespace flou
