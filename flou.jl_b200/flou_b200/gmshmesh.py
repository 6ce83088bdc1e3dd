"""UnstructuredMesh{2,Float64}(filename): MSH 4.1 (ASCII) quad meshes without libgmsh, and
UnstructuredMesh{3,Float64} from hexahedral tables (`RawHexMesh`; `_facemap_3d`, GmshMesh.jl:311-393,
face orientations 0..7).

Mirrors `src/FlouCommon/GmshMesh.jl:37-172` and `_facemap_2d` (:254-309).  In the reference
every id comes from libgmsh (Gmsh.jl v0.2.2 / gmsh_jll v4.10.2, third-party, not vendored):
`get_nodes`, `get_elements`, `create_edges` / `get_edges`, physical groups.  libgmsh is not
available here, so its numbering is RESTATED as a rule (parity unpinned -- no
reference test loads a mesh):

  * nodes are addressed by tag (the fixtures' tags are 1..N in file order);
  * elements of the top dimension are numbered in entity-block / file order;
  * `create_edges()` walks entities in (dim, tag) order, elements in file order and local
    edges in gmsh order (quad: (v0,v1), (v1,v2), (v2,v3), (v3,v0)), giving each NEW edge the
    next tag: boundary line elements come first, so an edge on the boundary has the tag of
    its line element -- which is what Flou relies on when it uses line-element tags as face
    ids (GmshMesh.jl:113-125).  `load` verifies that precondition and raises otherwise.

Everything Flou computes itself is followed exactly: local face order per quad = gmsh edges
[4, 2, 1, 3] (left, right, bottom, top; :263-268), node-order fix-up (:273-278), master = the
first element that lists the edge, orientation 0/1 from the first-node match (:297-306),
facepos/elempos by position search (:129-149), boundaries from the physical groups of
dimension ND-1 in group order with per-entity sorted line tags (:103-126).
"""
import numpy as np


class RawMesh:
    """What libgmsh would hand to Flou: nodes by tag, top-dimension elements in order, the
    boundary line elements by tag with their entity, and the physical groups."""

    def __init__(self, nodes, quads, lines, line_tags, line_entity, groups):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)        # (N, 2), row i = tag i+1
        self.quads = np.ascontiguousarray(quads, dtype=np.int64)          # (ne, 4) node tags, gmsh order
        self.lines = np.ascontiguousarray(lines, dtype=np.int64)          # (nb, 2) node tags
        self.line_tags = np.ascontiguousarray(line_tags, dtype=np.int64)  # (nb,) element tags
        self.line_entity = np.ascontiguousarray(line_entity, dtype=np.int64)
        self.groups = groups            # [(name, [entity tags])] in physical-group order


def read_msh(filename):
    """Minimal MSH 4.1 ASCII reader (sections PhysicalNames, Entities, Nodes, Elements)."""
    with open(filename) as fh:
        tok = fh.read().split("\n")
    sections, i = {}, 0
    while i < len(tok):
        line = tok[i].strip()
        if line.startswith("$") and not line.startswith("$End"):
            name, j = line[1:], i + 1
            while tok[j].strip() != "$End" + name:
                j += 1
            sections[name] = tok[i + 1:j]
            i = j
        i += 1
    fmt = sections["MeshFormat"][0].split()
    if not fmt[0].startswith("4.") or fmt[1] != "0":
        raise ValueError("only MSH 4.x ASCII files are supported")
    names = {}
    for ln in sections.get("PhysicalNames", [])[1:]:
        d, t, nm = ln.split(maxsplit=2)
        names[(int(d), int(t))] = nm.strip().strip('"')
    ent = sections["Entities"]
    npnt, ncur, nsur, nvol = (int(v) for v in ent[0].split())
    curve_phys = {}
    for ln in ent[1 + npnt:1 + npnt + ncur]:
        f = ln.split()
        nphys = int(f[7])
        curve_phys[int(f[0])] = [int(v) for v in f[8:8 + nphys]]
    # nodes
    nd_lines = sections["Nodes"]
    nblocks, nnodes, _, maxtag = (int(v) for v in nd_lines[0].split())
    coords = np.full((maxtag, 3), np.nan)
    p = 1
    for _ in range(nblocks):
        _, _, _, nb = (int(v) for v in nd_lines[p].split())
        tags = [int(nd_lines[p + 1 + q]) for q in range(nb)]
        for q, t in enumerate(tags):
            coords[t - 1] = [float(v) for v in nd_lines[p + 1 + nb + q].split()]
        p += 1 + 2 * nb
    if np.isnan(coords).any() or nnodes != maxtag:
        raise ValueError("node tags must be 1..N (as in the reference's fixtures)")
    # elements
    el = sections["Elements"]
    nblocks = int(el[0].split()[0])
    quads, quad_tags, lines, line_tags, line_entity = [], [], [], [], []
    p = 1
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(v) for v in el[p].split())
        for q in range(nb):
            f = [int(v) for v in el[p + 1 + q].split()]
            if edim == 2:
                if etype != 3:
                    raise ValueError("In 2D, all elements must be quadrilaterals.")
                quads.append(f[1:5]); quad_tags.append(f[0])
            elif edim == 1:
                if etype != 1:
                    raise ValueError("boundary elements must be 2-node lines")
                lines.append(f[1:3]); line_tags.append(f[0]); line_entity.append(etag)
        p += 1 + nb
    groups = []
    for (d, t) in sorted(k for k in names if k[0] == 1):
        ents = sorted(c for c, ph in curve_phys.items() if t in ph)
        groups.append((names[(d, t)], ents))
    order = np.lexsort((np.arange(len(lines)), np.array(line_entity)))    # entity, then file order
    return RawMesh(coords[:, :2], quads, np.array(lines).reshape(-1, 2)[order],
                   np.array(line_tags)[order], np.array(line_entity)[order], groups)


def refine(raw, r):
    """Synthetic r x r refinement of every quad by bilinear subdivision (config 5 scaling).
    Vertices on shared edges are merged; boundary lines are split into r pieces and renumbered
    1..r*nb in (entity, position) order so the boundary-tag rule keeps holding."""
    if r == 1:
        return raw
    key2id, nodes = {}, []

    def node_id(x):
        k = (round(float(x[0]), 12), round(float(x[1]), 12))
        if k not in key2id:
            key2id[k] = len(nodes) + 1
            nodes.append([float(x[0]), float(x[1])])
        return key2id[k]
    for x in raw.nodes:                     # keep the original vertices first (same tags)
        node_id(x)
    t = np.linspace(0.0, 1.0, r + 1)
    quads = []
    for q in raw.quads:
        v = raw.nodes[q - 1]
        ids = np.empty((r + 1, r + 1), dtype=np.int64)
        for j in range(r + 1):
            for i in range(r + 1):
                a, b = t[i], t[j]
                x = (v[0] * (1 - a) * (1 - b) + v[1] * a * (1 - b) + v[2] * a * b + v[3] * (1 - a) * b)
                ids[i, j] = node_id(x)
        for j in range(r):
            for i in range(r):
                quads.append([ids[i, j], ids[i + 1, j], ids[i + 1, j + 1], ids[i, j + 1]])
    lines, ent = [], []
    for (n1, n2), e in zip(raw.lines, raw.line_entity):
        a, b = raw.nodes[n1 - 1], raw.nodes[n2 - 1]
        ids = [node_id(a * (1 - s) + b * s) for s in t]
        for i in range(r):
            lines.append([ids[i], ids[i + 1]]); ent.append(e)
    return RawMesh(np.array(nodes), quads, lines, np.arange(1, len(lines) + 1), ent, raw.groups)


def write_msh(raw, filename):
    """Write a RawMesh as MSH 4.1 ASCII (one curve entity per boundary entity, one surface)."""
    ents = sorted(set(int(e) for e in raw.line_entity))
    phys_of = {e: [gi + 1 for gi, (_, es) in enumerate(raw.groups) if e in es] for e in ents}
    out = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat", "$PhysicalNames", str(len(raw.groups) + 1)]
    out += [f'1 {gi + 1} "{name}"' for gi, (name, _) in enumerate(raw.groups)]
    out += [f'2 {len(raw.groups) + 1} "Domain"', "$EndPhysicalNames", "$Entities",
            f"0 {len(ents)} 1 0"]
    for e in ents:
        ph = phys_of[e]
        out.append(f"{e} 0 0 0 0 0 0 {len(ph)} " + " ".join(map(str, ph)) + " 0")
    out.append(f"1 0 0 0 0 0 0 1 {len(raw.groups) + 1} 0")
    out += ["$EndEntities", "$Nodes", f"1 {len(raw.nodes)} 1 {len(raw.nodes)}",
            f"2 1 0 {len(raw.nodes)}"]
    out += [str(i + 1) for i in range(len(raw.nodes))]
    out += [f"{x[0]!r} {x[1]!r} 0" for x in raw.nodes.tolist()]
    out += ["$EndNodes", "$Elements",
            f"{len(ents) + 1} {len(raw.lines) + len(raw.quads)} 1 {len(raw.lines) + len(raw.quads)}"]
    for e in ents:
        sel = np.nonzero(raw.line_entity == e)[0]
        out.append(f"1 {e} 1 {len(sel)}")
        out += [f"{int(raw.line_tags[i])} {int(raw.lines[i, 0])} {int(raw.lines[i, 1])}" for i in sel]
    out.append(f"2 1 3 {len(raw.quads)}")
    nb = len(raw.lines)
    out += [f"{nb + i + 1} " + " ".join(str(int(v)) for v in q) for i, q in enumerate(raw.quads)]
    out += ["$EndElements", ""]
    with open(filename, "w") as fh:
        fh.write("\n".join(out))


class RawHexMesh:
    """What libgmsh would hand to Flou for a hexahedral mesh: nodes by tag, hexes in order (gmsh
    vertex order: 0-3 bottom counter-clockwise, 4-7 top), the boundary quads -- element tags
    1..Nb in the order given -- with their surface entity, and the physical groups."""

    def __init__(self, nodes, hexes, quads, quad_entity, groups):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)        # (N, 3), row i = tag i+1
        self.hexes = np.ascontiguousarray(hexes, dtype=np.int64)          # (ne, 8)
        self.quads = np.ascontiguousarray(quads, dtype=np.int64).reshape(-1, 4)
        self.quad_entity = np.ascontiguousarray(quad_entity, dtype=np.int64)
        self.groups = groups


# gmsh's local faces of a hexahedron (0-based vertices, MHexahedron order) re-ordered to Flou's
# local faces [xi-, xi+, eta-, eta+, zeta-, zeta+] = gmsh faces [3, 4, 2, 5, 1, 6] (GmshMesh.jl:320-329),
# each already permuted by `nodemap` (:333-340) so that the four corners run in the face-dof order of
# that local face (first tangential index fastest)
_HEX_FACE_CORNERS = np.array([[0, 3, 7, 4], [1, 2, 6, 5], [0, 1, 5, 4], [3, 2, 6, 7], [0, 1, 2, 3], [4, 5, 6, 7]])
# gmsh visits the local faces in ITS order when it hands out new face tags
_GMSH_VISIT = [4, 2, 0, 1, 3, 5]          # Flou local faces in gmsh order 1..6
# orientation code from the positions (0-based) of the master's first two corners in the slave's list
_ORIENT = {(0, 1): 0, (0, 3): 4, (1, 2): 1, (1, 0): 5, (2, 3): 2, (2, 1): 6, (3, 0): 3, (3, 2): 7}


class _HexMesh:
    """UnstructuredMesh(3, RawHexMesh): the tables of `UnstructuredMesh{3,Float64}` (GmshMesh.jl:49-172
    with `_facemap_3d`, :311-393).  Face ids follow the same rule as in 2-D: boundary quads keep
    their element tags 1..Nb, interior faces are numbered by first appearance walking the hexes in
    order and their faces in gmsh's local order (create_faces(); parity unpinned)."""

    cartesian = False
    nd = 3

    def __init__(self, raw):
        self.raw, self.nodes, self.nodeinds = raw, raw.nodes, raw.hexes
        ne, nb = len(raw.hexes), len(raw.quads)
        corners = raw.hexes[:, _HEX_FACE_CORNERS]                       # (ne, 6, 4) node tags
        keys = np.sort(corners, axis=2)
        tags = {tuple(k): i + 1 for i, k in enumerate(np.sort(raw.quads, axis=1).tolist())}
        if len(tags) != nb:
            raise ValueError("duplicated boundary quad")
        faceinds = np.zeros((ne, 6), dtype=np.int64)
        owner = {}                                                       # tag -> [(element, local face)]
        for e in range(ne):
            for lf in _GMSH_VISIT:
                k = tuple(keys[e, lf].tolist())
                t = tags.get(k)
                if t is None:
                    t = tags[k] = len(tags) + 1
                faceinds[e, lf] = t
                owner.setdefault(t, []).append((e, lf))
        nf = len(tags)
        eleminds = np.zeros((nf, 2), dtype=np.int64)
        elempos = np.zeros((nf, 2), dtype=np.int64)
        orientation = np.zeros(nf, dtype=np.uint8)
        self.face_nodeinds = np.zeros((nf, 4), dtype=np.int64)
        for t in range(1, nf + 1):
            own = owner.get(t)
            if not own:
                raise ValueError("a boundary quad does not coincide with an element face")
            if len(own) > 2:
                raise ValueError("a face is shared by more than two elements")
            (em, lm) = own[0]
            eleminds[t - 1, 0], elempos[t - 1, 0] = em + 1, lm + 1
            first = corners[em, lm]
            self.face_nodeinds[t - 1] = first
            if len(own) == 2:
                (es, ls) = own[1]
                eleminds[t - 1, 1], elempos[t - 1, 1] = es + 1, ls + 1
                second = corners[es, ls].tolist()
                orientation[t - 1] = _ORIENT[(second.index(first[0]), second.index(first[1]))]
        self.faceinds, self.eleminds, self.elempos, self.orientation = faceinds, eleminds, elempos, orientation
        self.facepos = np.where(eleminds[faceinds - 1, 0] == np.arange(1, ne + 1)[:, None], 1, 2)
        # an element that meets itself across a face is not representable (and not produced by gmsh)
        interior = eleminds[:, 1] != 0
        self.intfaces = np.nonzero(interior)[0].astype(np.int64) + 1
        self.bdnames, self.bdfaces = [], []
        for name, ents in raw.groups:
            faces = []
            for ent in ents:
                faces.extend(sorted(int(t) + 1 for t in np.nonzero(raw.quad_entity == ent)[0]))
            self.bdnames.append(name)
            self.bdfaces.append(np.array(faces, dtype=np.int64))
        self.bdmap = {i: i for i in range(1, len(self.bdfaces) + 1)}
        self.periodic = {}
        listed = np.concatenate(self.bdfaces) if self.bdfaces else np.zeros(0, dtype=np.int64)
        if not np.array_equal(np.sort(listed), np.nonzero(~interior)[0] + 1):
            raise ValueError("every boundary face must belong to exactly one physical group")

    @property
    def nelements(self):
        return self.faceinds.shape[0]

    @property
    def nfaces(self):
        return self.eleminds.shape[0]

    def nboundaries(self):
        return len(self.bdfaces)

    def element_vertices(self):
        return self.nodes[self.nodeinds - 1]


class UnstructuredMesh:
    """UnstructuredMesh(2, filename) == UnstructuredMesh{2,Float64}(filename);
    UnstructuredMesh(3, RawHexMesh) builds the 3-D tables (no 3-D .msh reader on this path)."""

    cartesian = False

    def __new__(cls, nd, source=None, refinement=1):
        if nd == 3 and isinstance(source, RawHexMesh):
            return _HexMesh(source)
        return super().__new__(cls)

    def __init__(self, nd, source, refinement=1):
        if nd != 2:
            raise ValueError("only 2-D quadrilateral .msh files are imported on this path "
                             "(3-D: pass a RawHexMesh)")
        raw = read_msh(source) if isinstance(source, str) else source
        raw = refine(raw, refinement)
        self.nd = 2
        self.raw = raw
        self.nodes = raw.nodes
        self.nodeinds = raw.quads
        ne, nb = len(raw.quads), len(raw.lines)
        if nb and not np.array_equal(raw.line_tags, np.arange(1, nb + 1)):
            raise ValueError("boundary line tags must be 1..Nb in entity order: Flou uses them as "
                             "face ids (GmshMesh.jl:113-125)")
        # ---- create_edges(): tags by first appearance, boundary lines first
        tags = {}
        for i, (a, b) in enumerate(raw.lines):
            tags[(min(a, b), max(a, b))] = i + 1
        if len(tags) != nb:
            raise ValueError("duplicated boundary line")
        q = raw.quads
        edge_nodes = np.stack([q[:, [0, 1]], q[:, [1, 2]], q[:, [2, 3]], q[:, [3, 0]]], axis=1)  # (ne,4,2)
        etags = np.empty((ne, 4), dtype=np.int64)
        for e in range(ne):
            for k in range(4):
                a, b = edge_nodes[e, k]
                key = (a, b) if a < b else (b, a)
                t = tags.get(key)
                if t is None:
                    t = tags[key] = len(tags) + 1
                etags[e, k] = t
        nf = len(tags)
        # ---- Flou's face order and node fix-up (GmshMesh.jl:263-278)
        self.faceinds = etags[:, [3, 1, 0, 2]].copy()
        nodemap = ([0, 1], [0, 1], [1, 0], [1, 0])
        eleminds = np.zeros((nf, 2), dtype=np.int64)
        first_nodes = np.zeros((nf, 2), dtype=np.int64)
        second_nodes = np.zeros((nf, 2), dtype=np.int64)
        for e in range(ne):
            for k in range(4):
                t = etags[e, k] - 1
                nodes = edge_nodes[e, k][nodemap[k]]
                if eleminds[t, 0] == 0:
                    eleminds[t, 0] = e + 1
                    first_nodes[t] = nodes
                else:
                    if eleminds[t, 1] != 0:
                        raise ValueError("an edge is shared by more than two elements")
                    eleminds[t, 1] = e + 1
                    second_nodes[t] = nodes
        if (eleminds[:, 0] == 0).any():
            raise ValueError("a boundary line does not coincide with an element edge")
        interior = eleminds[:, 1] != 0
        self.orientation = np.where(interior & (second_nodes[:, 0] != first_nodes[:, 0]), 1, 0).astype(np.uint8)
        self.eleminds = eleminds
        self.face_nodeinds = first_nodes
        # ---- facepos / elempos by position search (GmshMesh.jl:129-149)
        self.facepos = np.where(eleminds[self.faceinds - 1, 0] == np.arange(1, ne + 1)[:, None], 1, 2)
        elempos = np.zeros((nf, 2), dtype=np.int64)
        for s in range(2):
            has = eleminds[:, s] != 0
            fe = self.faceinds[eleminds[has, s] - 1]                   # (n, 4)
            elempos[has, s] = np.argmax(fe == (np.nonzero(has)[0] + 1)[:, None], axis=1) + 1
        self.elempos = elempos
        self.intfaces = np.nonzero(interior)[0].astype(np.int64) + 1
        # ---- boundaries from the physical groups (GmshMesh.jl:103-126)
        self.bdnames, self.bdfaces = [], []
        for name, ents in raw.groups:
            faces = []
            for ent in ents:
                faces.extend(sorted(int(t) for t in raw.line_tags[raw.line_entity == ent]))
            self.bdnames.append(name)
            self.bdfaces.append(np.array(faces, dtype=np.int64))
        self.bdmap = {i: i for i in range(1, len(self.bdfaces) + 1)}
        self.periodic = {}
        listed = np.concatenate(self.bdfaces) if self.bdfaces else np.zeros(0, dtype=np.int64)
        if not np.array_equal(np.sort(listed), np.nonzero(~interior)[0] + 1):
            raise ValueError("every boundary edge must belong to exactly one physical group")

    @property
    def nelements(self):
        return self.faceinds.shape[0]

    @property
    def nfaces(self):
        return self.eleminds.shape[0]

    def nboundaries(self):
        return len(self.bdfaces)

    def element_vertices(self):
        return self.nodes[self.nodeinds - 1]
