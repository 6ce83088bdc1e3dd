/* Prints the byte offset of every field of flou_b200_desc (include/flou_b200.h) as
 *   name offset size
 * lines, then "sizeof <bytes>".  tests/test_abi.py compiles and runs it (gcc, no GPU) and checks
 * the Python mirror (flou_b200/_lib.py: Desc) and the committed Julia table
 * (flou.jl_b200/julia/desc_offsets.jl, asserted by FlouB200.jl when the module loads) against it,
 * so that neither hand-written mirror can drift from the header silently. */
#include <stddef.h>
#include <stdio.h>
#include "flou_b200.h"

#define F(name) printf("%s %zu %zu\n", #name, offsetof(flou_b200_desc, name), sizeof(((flou_b200_desc *)0)->name))

int main(void)
{
    F(struct_size); F(nd); F(nv); F(np); F(equation); F(divop); F(tpflux); F(numflux);
    F(numflux_avg); F(geometry); F(intensity); F(gamma); F(a); F(dx); F(ne); F(nf);
    F(faceinds); F(facepos); F(eleminds); F(elempos); F(orientation);
    F(D); F(Ds); F(Dsharp); F(lminus); F(lplus); F(dgminus); F(dgplus); F(weights);
    F(jac); F(metric); F(fjac); F(frames);
    F(nbound); F(bc_kind); F(bc_offsets); F(bc_faces); F(bc_state); F(bc_table);
    F(elem_begin); F(elem_end); F(rank); F(nranks); F(part_offsets); F(device); F(flags);
    F(blend); F(sub_frames); F(sub_jac);
    printf("sizeof %zu\n", sizeof(flou_b200_desc));
    return 0;
}
