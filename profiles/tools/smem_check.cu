// Development check (host only, no GPU needed):  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17
//   -Iflou.jl_b200/csrc -Iinclude -o /tmp/smem_check profiles/tools/smem_check.cu && /tmp/smem_check
// lists every kernel instance whose dynamic shared memory exceeds the 227 KB a B200 SM offers per CTA.
#include <cstdio>
#include "launch.h"
#include "face_kernel.cuh"
#include "line_kernel.cuh"
using namespace flou;
template <int ND, int NP, int EQ, int VOL, bool CART, bool NB>
void line(const char *name) {
    using L = LCfg<ND, NP, EQ, VOL, CART, true, NB>;
    const double kb = L::SMEM_BYTES / 1024.0;
    if (kb > 227.0 || L::T > 1024)
        printf("line kernel  %-18s nd=%d np=%d cart=%d nb=%d: E=%d T=%d smem=%.1f KB\n", name, ND, NP, (int)CART, (int)NB, L::E, L::T, kb);
}
template <int ND, int NP, int EQ, int VOL, bool CART>
void stage(const char *name) {
    using K = KCfg<ND, NP, EQ, VOL, CART>;
    const double kb = K::SMEM_BYTES / 1024.0;
    if (kb > 227.0 || K::THREADS > 1024)
        printf("stage kernel %-18s nd=%d np=%d cart=%d: EPB=%d T=%d smem=%.1f KB\n", name, ND, NP, (int)CART, K::EPB, K::THREADS, kb);
}
template <int ND, int NP> void pair() {
    line<ND, NP, EQ_ADV, VOL_STRONG, true, false>("adv strong"); line<ND, NP, EQ_ADV, VOL_STRONG, false, false>("adv strong");
    line<ND, NP, EQ_ADV, VOL_SPLIT_STD, true, false>("adv split"); line<ND, NP, EQ_ADV, VOL_SPLIT_STD, false, false>("adv split");
    line<ND, NP, EQ_EULER, VOL_STRONG, true, false>("euler strong"); line<ND, NP, EQ_EULER, VOL_STRONG, false, false>("euler strong");
    line<ND, NP, EQ_EULER, VOL_SPLIT_STD, true, false>("euler split-std"); line<ND, NP, EQ_EULER, VOL_SPLIT_STD, false, false>("euler split-std");
    line<ND, NP, EQ_EULER, VOL_SPLIT_CHA, true, false>("euler split-cha"); line<ND, NP, EQ_EULER, VOL_SPLIT_CHA, false, false>("euler split-cha");
    line<ND, NP, EQ_EULER, VOL_HYBRID, true, false>("euler hybrid"); line<ND, NP, EQ_EULER, VOL_HYBRID, false, false>("euler hybrid");
    line<ND, NP, EQ_EULER, VOL_SPLIT_STD, true, true>("euler split-std nb"); line<ND, NP, EQ_EULER, VOL_SPLIT_STD, false, true>("euler split-std nb");
    line<ND, NP, EQ_EULER, VOL_SPLIT_CHA, true, true>("euler split-cha nb"); line<ND, NP, EQ_EULER, VOL_SPLIT_CHA, false, true>("euler split-cha nb");
    line<ND, NP, EQ_EULER, VOL_HYBRID, true, true>("euler hybrid nb"); line<ND, NP, EQ_EULER, VOL_HYBRID, false, true>("euler hybrid nb");
    stage<ND, NP, EQ_ADV, VOL_STRONG, true>("adv strong"); stage<ND, NP, EQ_ADV, VOL_STRONG, false>("adv strong");
    stage<ND, NP, EQ_ADV, VOL_SPLIT_STD, true>("adv split"); stage<ND, NP, EQ_ADV, VOL_SPLIT_STD, false>("adv split");
    stage<ND, NP, EQ_EULER, VOL_STRONG, true>("euler strong"); stage<ND, NP, EQ_EULER, VOL_STRONG, false>("euler strong");
    stage<ND, NP, EQ_EULER, VOL_SPLIT_STD, true>("euler split-std"); stage<ND, NP, EQ_EULER, VOL_SPLIT_STD, false>("euler split-std");
    stage<ND, NP, EQ_EULER, VOL_SPLIT_CHA, true>("euler split-cha"); stage<ND, NP, EQ_EULER, VOL_SPLIT_CHA, false>("euler split-cha");
}
int main() {
    pair<1,2>(); pair<1,3>(); pair<1,4>(); pair<1,5>(); pair<1,6>(); pair<1,7>(); pair<1,8>();
    pair<2,2>(); pair<2,3>(); pair<2,4>(); pair<2,5>(); pair<2,6>(); pair<2,7>(); pair<2,8>();
    pair<3,2>(); pair<3,3>(); pair<3,4>(); pair<3,5>(); pair<3,6>(); pair<3,7>(); pair<3,8>();
    printf("checked 21 (nd, np) pairs\n");
}
