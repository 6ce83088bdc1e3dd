"""ORACLE (test infrastructure, never shipped, never on the product path).

Literal restatement of `UnstructuredMesh{2,Float64}(filename)` -- src/FlouCommon/GmshMesh.jl:37-172
and `_facemap_2d` :254-309 -- with the libgmsh calls (Gmsh.jl v0.2.2 / gmsh_jll v4.10.2, third
party, absent) replaced by the numbering rule of SURVEY.md 8(c):
`create_edges()` numbers edges by first appearance walking entities in (dim, tag) order,
elements in file order, local quad edges (v0,v1), (v1,v2), (v2,v3), (v3,v0).  Parity unpinned:
no reference test loads a mesh.  Dict-for-dict like the Julia source; small meshes only.
"""
import numpy as np

from .connectivity import Mesh


def parse_msh41(filename):
    """-> nodes {tag: (x,y,z)}, blocks [(dim, entity, type, [(tag, nodes...)])],
    names {(dim,tag): name}, curve_phys {curve: [phys tags]}."""
    lines = [ln.strip() for ln in open(filename)]
    pos = {ln: i for i, ln in enumerate(lines) if ln.startswith("$")}
    names = {}
    if "$PhysicalNames" in pos:
        i = pos["$PhysicalNames"] + 1
        for ln in lines[i + 1:i + 1 + int(lines[i])]:
            d, t, nm = ln.split(maxsplit=2)
            names[(int(d), int(t))] = nm.strip('"')
    i = pos["$Entities"] + 1
    npnt, ncur = (int(v) for v in lines[i].split()[:2])
    curve_phys = {}
    for ln in lines[i + 1 + npnt:i + 1 + npnt + ncur]:
        f = ln.split()
        curve_phys[int(f[0])] = [int(v) for v in f[8:8 + int(f[7])]]
    i = pos["$Nodes"] + 1
    nblocks = int(lines[i].split()[0])
    i += 1
    nodes = {}
    for _ in range(nblocks):
        nb = int(lines[i].split()[3])
        tags = [int(lines[i + 1 + q]) for q in range(nb)]
        for q, t in enumerate(tags):
            nodes[t] = tuple(float(v) for v in lines[i + 1 + nb + q].split())
        i += 1 + 2 * nb
    i = pos["$Elements"] + 1
    nblocks = int(lines[i].split()[0])
    i += 1
    blocks = []
    for _ in range(nblocks):
        dim, ent, typ, nb = (int(v) for v in lines[i].split())
        blocks.append((dim, ent, typ, [tuple(int(v) for v in lines[i + 1 + q].split()) for q in range(nb)]))
        i += 1 + nb
    return nodes, blocks, names, curve_phys


def unstructured_mesh_2d(filename=None, parsed=None):
    nodes_by_tag, blocks, names, curve_phys = parsed if parsed is not None else parse_msh41(filename)
    # get_nodes(): Flou indexes `nodes` by tag (GmshMesh.jl:51-58, 77-83)
    numnodes = len(nodes_by_tag)
    nodes = np.array([nodes_by_tag[t][:2] for t in range(1, numnodes + 1)])
    # get_elements(2): entity order, file order; a single element type
    quad_blocks = sorted((b for b in blocks if b[0] == 2), key=lambda b: b[1])
    if len({b[2] for b in quad_blocks}) != 1:
        raise ValueError("Hybrid meshes are not supported.")
    if quad_blocks[0][2] != 3:
        raise ValueError("In 2D, all elements must be quadrilaterals.")
    elemnodes = [list(el[1:5]) for b in quad_blocks for el in b[3]]
    numelements = len(elemnodes)
    # create_edges(): the numbering rule
    edge_tag = {}
    for b in sorted((b for b in blocks if b[0] == 1), key=lambda b: b[1]):
        for el in b[3]:
            key = frozenset(el[1:3])
            if key not in edge_tag:
                edge_tag[key] = len(edge_tag) + 1
            if edge_tag[key] != el[0]:
                raise ValueError("line-element tag differs from its edge tag (GmshMesh.jl:113-125)")
    ntags, etags = [], []           # get_element_edge_nodes / get_edges
    for en in elemnodes:
        for a, b in ((0, 1), (1, 2), (2, 3), (3, 0)):
            ntags += [en[a], en[b]]
            key = frozenset((en[a], en[b]))
            if key not in edge_tag:
                edge_tag[key] = len(edge_tag) + 1
            etags.append(edge_tag[key])
    # _facemap_2d (GmshMesh.jl:254-309), 1-based arithmetic kept
    element2edge = {}
    cnt = 0
    for i in range(1, numelements + 1):
        element2edge[i] = [etags[cnt + 4 - 1], etags[cnt + 2 - 1], etags[cnt + 1 - 1], etags[cnt + 3 - 1]]
        cnt += 4
    nodemap = ((1, 2), (1, 2), (2, 1), (2, 1))
    edge2element, edge2node, e2n_second = {}, {}, {}
    cnt = 0
    for i, etag in enumerate(etags, start=1):
        ielem = (i - 1) // 4 + 1
        pos = (i - 1) % 4 + 1
        pair = [ntags[cnt + nodemap[pos - 1][0] - 1], ntags[cnt + nodemap[pos - 1][1] - 1]]
        if etag in edge2element:
            edge2element[etag][1] = ielem
            e2n_second[etag] = pair
        else:
            edge2element[etag] = [ielem, 0]
            edge2node[etag] = pair
        cnt += 2
    orientations = {t: 0 for t in edge2node}
    for iedge, nodes2 in e2n_second.items():
        node = edge2node[iedge][0]
        orientations[iedge] = 0 if nodes2.index(node) == 0 else 1
    faceinds, eleminds = element2edge, edge2element
    numfaces = len(eleminds)
    intfaces = sorted(i for i, f in eleminds.items() if f[1] != 0)
    # boundaries (GmshMesh.jl:103-126)
    bdnames, bdfaces = [], []
    for (dim, tag) in sorted(k for k in names if k[0] == 1):
        bdnames.append(names[(dim, tag)])
        faces = []
        for entity in sorted(c for c, ph in curve_phys.items() if tag in ph):
            ftags = sorted(el[0] for b in blocks if b[0] == 1 and b[1] == entity for el in b[3])
            faces += ftags
        bdfaces.append(faces)
    # elements and faces (GmshMesh.jl:127-149)
    facepos = [[eleminds[f].index(i) + 1 for f in faceinds[i]] for i in range(1, numelements + 1)]
    elempos = []
    for i in range(1, numfaces + 1):
        elempos.append([0 if ie == 0 else faceinds[ie].index(i) + 1 for ie in eleminds[i]])
    mesh = Mesh(2, (), (), nodes, [faceinds[i] for i in range(1, numelements + 1)], facepos,
                [list(eleminds[i]) for i in range(1, numfaces + 1)], elempos,
                [orientations[i] for i in range(1, numfaces + 1)], intfaces, bdfaces, bdnames,
                {i: i for i in range(1, len(bdfaces) + 1)})
    mesh.enodes = elemnodes
    return mesh


# gmsh's local faces of an 8-node hexahedron (MHexahedron, 0-based vertex numbers), the order in
# which get_element_face_nodes(etype, 4) lists them
_GMSH_HEX_FACES = ((0, 3, 2, 1), (0, 1, 5, 4), (0, 4, 7, 3), (1, 2, 6, 5), (2, 3, 7, 6), (4, 5, 6, 7))


def unstructured_mesh_3d(nodes, hexes, quads, quad_entity, groups):
    """Literal restatement of `UnstructuredMesh{3,Float64}()` (GmshMesh.jl:49-172) and
    `_facemap_3d` (:311-393) for hexahedral meshes given as tables: `nodes` (N, 3) by tag, `hexes`
    node tags in gmsh order, boundary `quads` (tags 1..Nb in the order given) with their surface
    entity, `groups` [(name, [entities])].  libgmsh's `create_faces()` numbering is restated by the
    same rule as in 2-D: boundary elements first, then element faces by first appearance (gmsh
    local face order).  Parity unpinned (no reference test loads a mesh)."""
    nodes = np.asarray(nodes, dtype=float)
    elemnodes = [list(map(int, h)) for h in hexes]
    numelements = len(elemnodes)
    face_tag = {}
    for q in quads:
        key = frozenset(int(v) for v in q)
        if key in face_tag:
            raise ValueError("duplicated boundary quad")
        face_tag[key] = len(face_tag) + 1
    ntags, ftags = [], []               # get_element_face_nodes(etype, 4) / get_faces(4, ntags)
    for en in elemnodes:
        for loc in _GMSH_HEX_FACES:
            fn = [en[v] for v in loc]
            ntags += fn
            key = frozenset(fn)
            if key not in face_tag:
                face_tag[key] = len(face_tag) + 1
            ftags.append(face_tag[key])
    # _facemap_3d, 1-based arithmetic kept
    element2face = {}
    cnt = 0
    for i in range(1, numelements + 1):
        element2face[i] = [ftags[cnt + 3 - 1], ftags[cnt + 4 - 1], ftags[cnt + 2 - 1],
                           ftags[cnt + 5 - 1], ftags[cnt + 1 - 1], ftags[cnt + 6 - 1]]
        cnt += 6
    nodemap = ((1, 4, 3, 2), (1, 2, 3, 4), (1, 4, 3, 2), (1, 2, 3, 4), (2, 1, 4, 3), (1, 2, 3, 4))
    face2element, face2node, f2n_second = {}, {}, {}
    cnt = 0
    for i, ftag in enumerate(ftags, start=1):
        ielem = (i - 1) // 6 + 1
        pos = (i - 1) % 6 + 1
        four = [ntags[cnt + m - 1] for m in nodemap[pos - 1]]
        if ftag in face2element:
            face2element[ftag][1] = ielem
            f2n_second[ftag] = four
        else:
            face2element[ftag] = [ielem, 0]
            face2node[ftag] = four
        cnt += 4
    orientations = {t: 0 for t in face2node}
    table = {(1, 2): 0, (1, 4): 4, (2, 3): 1, (2, 1): 5, (3, 4): 2, (3, 2): 6, (4, 1): 3, (4, 3): 7}
    for iface, nodes2 in f2n_second.items():
        n1, n2 = face2node[iface][0], face2node[iface][1]
        orientations[iface] = table[(nodes2.index(n1) + 1, nodes2.index(n2) + 1)]
    faceinds, eleminds = element2face, face2element
    numfaces = len(eleminds)
    intfaces = sorted(i for i, f in eleminds.items() if f[1] != 0)
    bdnames, bdfaces = [], []
    quad_entity = [int(e) for e in quad_entity]
    for name, ents in groups:
        bdnames.append(name)
        faces = []
        for entity in ents:
            faces += sorted(t + 1 for t, e in enumerate(quad_entity) if e == entity)
        bdfaces.append(faces)
    facepos = [[eleminds[f].index(i) + 1 for f in faceinds[i]] for i in range(1, numelements + 1)]
    elempos = []
    for i in range(1, numfaces + 1):
        elempos.append([0 if ie == 0 else faceinds[ie].index(i) + 1 for ie in eleminds[i]])
    mesh = Mesh(3, (), (), nodes, [faceinds[i] for i in range(1, numelements + 1)], facepos,
                [list(eleminds[i]) for i in range(1, numfaces + 1)], elempos,
                [orientations[i] for i in range(1, numfaces + 1)], intfaces, bdfaces, bdnames,
                {i: i for i in range(1, len(bdfaces) + 1)})
    mesh.enodes = elemnodes
    return mesh
