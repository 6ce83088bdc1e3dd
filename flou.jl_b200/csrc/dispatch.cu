// Stitches the per-(ND, NP) kernel tables (inst.cu) into one lookup.  The pair list comes
// from the Makefile through the generated build/pairs.h:  #define FLOU_PAIR_LIST X(1,2) ...
#include "launch.h"
#include "pairs.h"

namespace flou {

#define X(nd, np)                                                            \
    const StageLauncher *stage_table_##nd##_##np(int eq, int vol, int cart); \
    const EmitLauncher *emit_table_##nd##_##np(int nv);
FLOU_PAIR_LIST
#undef X

const StageLauncher *get_stage_launcher(int nd, int np, int eq, int vol, int cart)
{
    if (eq < 0 || eq > 1 || vol < 0 || vol > 6) return nullptr;
#define X(a, b) if (nd == a && np == b) return stage_table_##a##_##b(eq, vol, cart);
    FLOU_PAIR_LIST
#undef X
    return nullptr;
}

const EmitLauncher *get_emit_launcher(int nd, int np, int nv)
{
#define X(a, b) if (nd == a && np == b) return emit_table_##a##_##b(nv);
    FLOU_PAIR_LIST
#undef X
    return nullptr;
}

}  // namespace flou
