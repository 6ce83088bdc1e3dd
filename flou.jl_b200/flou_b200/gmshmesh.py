"""UnstructuredMesh{2,Float64}(filename): MSH 4.1 (ASCII) quad meshes without libgmsh.

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


class UnstructuredMesh:
    """UnstructuredMesh(2, filename) == UnstructuredMesh{2,Float64}(filename)."""

    cartesian = False

    def __init__(self, nd, source, refinement=1):
        if nd != 2:
            raise ValueError("only 2-D quadrilateral meshes are imported on this path")
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
