"""MultielementDisc / EquationConfig / rhs! on the B200 library.

  MultielementDisc(mesh, std, equation, operators, bcs)   src/FlouSpatial/MultielementDiscontinuous.jl:29-92
  EquationConfig(disc, equation)                          src/FlouCommon/FlouCommon.jl (struct used by rhs!)
  rhs!(dQ, Q, p, t)                                       src/FlouSpatial/Equations/Hyperbolic.jl:31-69
  GlobalStateVector layout (ndofs, nv) column-major       src/FlouSpatial/GlobalContainers.jl:19-29

The constructor packs the tables the reference's objects hold into a `flou_b200_desc` and
creates the device handle; nothing is computed on the host afterwards.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import geometry as G
from .equations import Frame, GenericBC, Source, _DependsOn
from .mesh import partition_offsets


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class EquationConfig:
    def __init__(self, disc, equation):
        self.disc, self.equation = disc, equation


class MultielementDisc:
    def __init__(self, mesh, std, equation, operators, bcs, source=None, *,
                 rank=0, nranks=1, device=None, geometry=None, use_graph=True, create=True,
                 fused=None, kernel=None):
        # source term (MultielementDiscontinuous.jl:75-79, 139-146): None = the default no-op;
        # a plain callable f(Q_i, x_i, t) -> increment of dQ_i is treated as depending on
        # everything (re-tabulated every stage); Source(f, state=False, time=False) is static
        if source is not None and not isinstance(source, Source):
            if not callable(source):
                raise ValueError("source must be a callable f(Q, x, t) or a flou_b200.Source")
            source = Source(source)
        self.source = source
        if std.nd != mesh.nd or equation.nd != mesh.nd:
            raise ValueError("mesh, standard region and equation dimensions differ")
        self.mesh, self.std, self.equation = mesh, std, equation
        self.operators = (operators,) if not isinstance(operators, (tuple, list)) else tuple(operators)
        op = self.operators[0]
        if op.kind == L.OP_SPLIT and not std.basis.hasboundaries and equation.kind != L.EQ_EULER:
            # _splitdiv_nb_surface_contribution! (OpDivergence.jl:300-437) needs vars_cons2entropy,
            # which the reference defines for the Euler equations only
            raise ValueError("SplitDivOperator on Gauss nodes needs entropy variables: Euler equations only")
        if op.kind == L.OP_HYBRID:
            if equation.kind != L.EQ_EULER:
                raise ValueError("HybridDivOperator needs entropy variables: Euler equations only")
        nd, npn, nv = mesh.nd, std.np, equation.nv
        self.nd, self.np, self.nv = nd, npn, nv
        self.npts, self.nfp = npn ** nd, npn ** (nd - 1)
        self.rank, self.nranks = int(rank), int(nranks)
        self.part_offsets = partition_offsets(mesh.nelements, self.nranks)
        self.elem_begin = int(self.part_offsets[self.rank])
        self.elem_end = int(self.part_offsets[self.rank + 1])
        self.ndofs_global = mesh.nelements * self.npts
        self.ndofs = (self.elem_end - self.elem_begin) * self.npts

        # ---- boundary conditions ordered by mesh.bdmap (MultielementDiscontinuous.jl:39-51)
        nb = mesh.nboundaries()
        if len(bcs) != nb:
            raise ValueError("The number of BCs does not match the number of boundaries.")
        ordered = [None] * nb
        if isinstance(bcs, dict):
            for key, value in bcs.items():
                j = mesh.bdnames.index(key) + 1
                ordered[mesh.bdmap[j] - 1] = value
        else:
            ordered = list(bcs)
        self.bcs = tuple(ordered)

        # ---- geometry
        cart = getattr(mesh, "cartesian", False) if geometry is None else (geometry == "cartesian")
        self.cartesian = cart
        self._verts = mesh.element_vertices()
        keep = self._keep = {}
        if not cart:
            jac, metric = G.general_element_geometry(self._verts, std.xi)
            _, fjac, frames = self._face_geometry()
            keep.update(jac=jac, metric=metric, fjac=fjac, frames=frames)
            if op.kind == L.OP_HYBRID or (op.kind == L.OP_SPLIT and not std.basis.hasboundaries):   # incl. hybrid on Gauss nodes
                # requires_subgrid(op, std): geometry.subgrids (PhysicalRegions.jl:179-292)
                sfr, sjac = G.general_subgrid_geometry(self._verts, std.xi1d, std.w1d)
                keep.update(sub_frames=sfr, sub_jac=sjac)

        # ---- descriptor
        d = self._desc = L.Desc()
        d.struct_size = C.sizeof(L.Desc)
        d.nd, d.nv, d.np = nd, nv, npn
        d.equation = equation.kind
        d.divop = op.kind
        d.tpflux = op.tpflux.kind if op.tpflux is not None else L.FLUX_STDAVERAGE
        d.blend = float(getattr(op, "blend", 0.0))
        nf_ = op.numflux
        d.numflux = nf_.kind
        d.numflux_avg = getattr(getattr(nf_, "avg", None), "kind", L.FLUX_STDAVERAGE)
        d.intensity = float(getattr(nf_, "intensity", 0.0))
        d.gamma = float(getattr(equation, "gamma", 0.0))
        for c in range(3):
            d.a[c] = equation.a[c] if hasattr(equation, "a") and c < nd else 0.0
            d.dx[c] = mesh.dx[c] if cart and c < nd else 0.0
        d.geometry = L.GEOM_CARTESIAN if cart else L.GEOM_GENERAL
        d.ne, d.nf = mesh.nelements, mesh.nfaces
        keep["faceinds"] = np.ascontiguousarray(mesh.faceinds, dtype=np.int64)
        keep["facepos"] = np.ascontiguousarray(mesh.facepos, dtype=np.int64)
        keep["eleminds"] = np.ascontiguousarray(mesh.eleminds, dtype=np.int64)
        keep["elempos"] = np.ascontiguousarray(mesh.elempos, dtype=np.int64)
        keep["orientation"] = np.ascontiguousarray(mesh.orientation, dtype=np.uint8)
        for name, arr in (("D", std.D), ("Ds", std.Ds), ("Dsharp", std.Dsharp)):
            keep[name] = np.ascontiguousarray(np.asarray(arr).T)   # column-major for the ABI
        keep["lminus"], keep["lplus"] = (np.ascontiguousarray(v) for v in std.l)
        keep["dgminus"], keep["dgplus"] = (np.ascontiguousarray(v) for v in std.dg)
        keep["weights"] = np.ascontiguousarray(std.w1d)
        kinds = np.array([bc.kind for bc in self.bcs] + [0], dtype=np.int32)
        offs = np.concatenate(([0], np.cumsum([len(b) for b in mesh.bdfaces]))).astype(np.int64)
        faces = (np.concatenate(list(mesh.bdfaces) + [np.zeros(1, dtype=np.int64)])
                 .astype(np.int64))
        state = np.zeros((max(nb, 1), nv))
        table = np.zeros((max(int(offs[-1]), 1) * self.nfp, nv))
        self._dynamic_bcs = []              # [(boundary index, GenericBC)] re-tabulated every stage
        if any(isinstance(bc, GenericBC) for bc in self.bcs):
            fcoords = self._fcoords = self.face_coords()
        for ib, bc in enumerate(self.bcs):
            if bc.kind == L.BC_INFLOW:
                if len(bc.Qext) != nv:
                    raise ValueError("EulerInflowBC state length does not match the equation")
                state[ib] = bc.Qext
            elif bc.kind == L.BC_TABLE:
                rows = range(int(offs[ib]), int(offs[ib + 1]))
                dynamic = bc.dynamic
                if dynamic is None and len(rows):      # probe: does the closure read Qin / frame / time?
                    try:
                        bc.tabulate(fcoords[(int(faces[rows[0]]) - 1) * self.nfp], equation)
                        dynamic = False
                    except _DependsOn:
                        dynamic = True
                if dynamic:
                    self._dynamic_bcs.append((ib, bc))
                    continue
                for m in rows:
                    f = int(faces[m]) - 1
                    for i in range(self.nfp):
                        table[m * self.nfp + i] = bc.tabulate(fcoords[f * self.nfp + i], equation)
        keep.update(bc_kind=kinds, bc_offsets=offs, bc_faces=faces,
                    bc_state=np.ascontiguousarray(state), bc_table=np.ascontiguousarray(table),
                    part_offsets=np.ascontiguousarray(self.part_offsets))
        for name in ("faceinds", "facepos", "eleminds", "elempos", "orientation", "D", "Ds",
                     "Dsharp", "lminus", "lplus", "dgminus", "dgplus", "weights", "bc_kind", "bc_offsets",
                     "bc_faces", "bc_state", "bc_table"):
            setattr(d, name, _ptr(keep[name]))
        if not cart:
            for name in ("jac", "metric", "fjac", "frames"):
                setattr(d, name, _ptr(keep[name]))
            if "sub_frames" in keep:
                d.sub_frames, d.sub_jac = _ptr(keep["sub_frames"]), _ptr(keep["sub_jac"])
        d.nbound = nb
        d.elem_begin, d.elem_end = self.elem_begin, self.elem_end
        d.rank, d.nranks = self.rank, self.nranks
        d.part_offsets = _ptr(keep["part_offsets"]) if self.nranks > 1 else None
        d.device = int(device if device is not None else 0)
        import os as _os
        if fused is None:
            fused = _os.environ.get("FLOU_B200_FUSED", "0") == "1"
        # kernel: None/"auto" (library heuristic: two-kernel stage with the line kernel, fused
        # single-kernel stage on meshes too small to fill the GPU), "line", "node" (two-kernel
        # stage with the node-per-thread element kernel), "fused"
        kernel = kernel or _os.environ.get("FLOU_B200_KERNEL", "auto")
        if _os.environ.get("FLOU_B200_NODE_KERNEL", "0") == "1":
            kernel = "node"
        if fused:
            kernel = "fused"
        if kernel not in ("auto", "line", "node", "fused"):
            raise ValueError(f"unknown kernel choice {kernel!r}")
        self.kernel = kernel
        d.flags = ((0 if use_graph else L.FLAG_NO_GRAPH)
                   | {"auto": 0, "line": L.FLAG_LINE_KERNEL, "node": L.FLAG_NODE_KERNEL,
                      "fused": L.FLAG_FUSED}[kernel])
        self._h = C.c_void_p()
        if create:
            L.check(L.lib().flou_b200_create(C.byref(d), C.byref(self._h)))
            # the library copied everything it needs; drop the big host tables
            for name in ("jac", "metric", "fjac", "frames", "sub_frames", "sub_jac"):
                keep.pop(name, None)
            self._init_dynamic()

    # ------------------------------------------------------------------ source / dynamic BCs
    @property
    def has_dynamic(self):
        """True when source or boundary data have to be re-tabulated by the host before every
        stage (the RK loop then runs stage by stage instead of from the captured graph)."""
        return bool(self._dynamic_bcs) or (self.source is not None and self.source.dynamic)

    def _init_dynamic(self):
        self._bc_table = self._keep["bc_table"]
        if self.source is not None:
            self._xloc = self.coords()[self.local_rows()]
            if not self.source.dynamic:
                S = self.source.tabulate(None, self._xloc, 0.0, self.nv)
                L.check(L.lib().flou_b200_set_source(self.handle, _ptr(S)))
        if self._dynamic_bcs:
            n = C.c_int64(0)
            L.check(L.lib().flou_b200_boundary_traces(self.handle, None, None, C.byref(n)))
            self._bd_ord = np.zeros(max(n.value, 1), dtype=np.int64)
            L.check(L.lib().flou_b200_boundary_traces(self.handle, None, _ptr(self._bd_ord), C.byref(n)))
            self._bd_ord = self._bd_ord[:n.value]
            self._bd_traces = np.zeros((max(n.value, 1), self.nv, self.nfp))
            offs = self._keep["bc_offsets"]
            self._bd_ib = np.searchsorted(offs, self._bd_ord, side="right") - 1      # boundary of each owned face
            _, _, frames = self._face_geometry()
            self._fframes = frames.reshape(-1, 3, self.nd)

    def refresh(self, t, Q=None):
        """Re-tabulate what depends on (Q, t) -- the dynamic source and GenericBC closures -- for a
        pass at time t on the device-resident state (`Q`: that state on the host, if the caller
        has it; downloaded otherwise when the source needs it)."""
        src = self.source
        if src is not None and src.dynamic:
            if src.state and Q is None:
                Q = self.download()
            S = src.tabulate(Q if src.state else None, self._xloc, float(t), self.nv)
            L.check(L.lib().flou_b200_set_source(self.handle, _ptr(S)))
        if self._dynamic_bcs:
            n = C.c_int64(0)
            L.check(L.lib().flou_b200_boundary_traces(self.handle, _ptr(self._bd_traces), None, C.byref(n)))
            dyn = dict(self._dynamic_bcs)
            faces = self._keep["bc_faces"]
            for j in range(n.value):
                bc = dyn.get(int(self._bd_ib[j]))
                if bc is None:
                    continue
                m = int(self._bd_ord[j])
                f = int(faces[m]) - 1
                for i in range(self.nfp):
                    fr = self._fframes[f * self.nfp + i]
                    self._bc_table[m * self.nfp + i] = bc.evaluate(
                        self._bd_traces[j, :, i].copy(), self._fcoords[f * self.nfp + i],
                        Frame(fr[0], fr[1], fr[2]), float(t), self.equation)
            L.check(L.lib().flou_b200_set_bc_table(self.handle, _ptr(self._bc_table)))

    def partition_plan(self):
        """Host-only halo plan (no GPU needed): dict with peers, per-peer slot counts, the
        global face id and (local element, local face) of every ghost slot."""
        lib = L.lib()
        ng, npeer = C.c_int64(0), C.c_int32(0)
        ni, nb = C.c_int64(0), C.c_int64(0)
        L.check(lib.flou_b200_partition_plan(C.byref(self._desc), C.byref(ng), C.byref(npeer),
                                             None, None, None, None, C.byref(ni), C.byref(nb)))
        peers = np.zeros(max(npeer.value, 1), dtype=np.int32)
        counts = np.zeros(max(npeer.value, 1), dtype=np.int64)
        faces = np.zeros(max(ng.value, 1), dtype=np.int64)
        elemfaces = np.zeros(max(ng.value, 1), dtype=np.int32)
        L.check(lib.flou_b200_partition_plan(C.byref(self._desc), C.byref(ng), C.byref(npeer),
                                             _ptr(peers), _ptr(counts), _ptr(faces),
                                             _ptr(elemfaces), C.byref(ni), C.byref(nb)))
        return dict(nghost=ng.value, peers=peers[:npeer.value].tolist(),
                    counts=counts[:npeer.value].tolist(), faces=faces[:ng.value],
                    elemfaces=elemfaces[:ng.value], n_interior=ni.value, n_boundary=nb.value)

    # ------------------------------------------------------------------ geometry access
    def _face_geometry(self):
        xif = self.std.xi[:self.nfp, :self.nd - 1] if self.nd > 1 else np.zeros((1, 0))
        if self.nd > 1:
            grids = np.meshgrid(*([self.std.xi1d] * (self.nd - 1)), indexing="ij")
            xif = np.stack([g.reshape(-1, order="F") for g in grids], axis=1)
        return G.general_face_geometry(self._verts, np.asarray(self.mesh.eleminds),
                                       np.asarray(self.mesh.elempos), xif, self.nd)

    def coords(self):
        """geometry.elements.coords: (ndofs_global, nd) node coordinates."""
        return G.element_coords(self._verts, self.std.xi)

    def face_coords(self):
        return self._face_geometry()[0]

    def local_rows(self):
        """Row range of this rank inside the global (ndofs, nv) state matrix."""
        return slice(self.elem_begin * self.npts, self.elem_end * self.npts)

    def new_state(self, local=True):
        return np.zeros((self.ndofs if local else self.ndofs_global, self.nv), order="F")

    # ------------------------------------------------------------------ device calls
    @property
    def handle(self):
        if not self._h:
            raise L.FlouB200Error("discretisation already closed")
        return self._h

    def upload(self, Q):
        Q = _state(Q, self.ndofs, self.nv)
        L.check(L.lib().flou_b200_upload_state(self.handle, _ptr(Q)))

    def download(self, out=None):
        out = self.new_state() if out is None else _state(out, self.ndofs, self.nv, writable=True)
        L.check(L.lib().flou_b200_download_state(self.handle, _ptr(out)))
        return out

    def get_max_dt(self, cfl, Q=None):
        """get_max_dt(q, disc, equation, cfl): global CFL time step (device reduction)."""
        dt = C.c_double(0.0)
        q = None if Q is None else _ptr(_state(Q, self.ndofs, self.nv))
        L.check(L.lib().flou_b200_max_dt(self.handle, q, float(cfl), C.byref(dt)))
        return dt.value

    def synchronize(self):
        L.check(L.lib().flou_b200_synchronize(self.handle))

    def status(self):
        f = C.c_int32(0)
        L.check(L.lib().flou_b200_status(self.handle, C.byref(f)))
        return f.value

    def kernel_launches(self):
        return int(L.lib().flou_b200_kernel_launches(self.handle))

    def kernel_info(self):
        v = [C.c_int32(0) for _ in range(4)]
        L.check(L.lib().flou_b200_kernel_info(self.handle, *[C.byref(x) for x in v]))
        return dict(grid_ctas=v[0].value, threads=v[1].value, smem_bytes=v[2].value,
                    elems_per_cta_iter=v[3].value)

    def profile(self, enable):
        """Per-kernel event timing: returns (ms in face kernel, ms in element kernel, passes)
        accumulated since the last call and switches the instrumentation on/off."""
        a, b, n = C.c_float(0), C.c_float(0), C.c_int64(0)
        L.check(L.lib().flou_b200_profile(self.handle, 1 if enable else 0, C.byref(a), C.byref(b),
                                          C.byref(n)))
        return a.value, b.value, n.value

    def timer_start(self):
        L.check(L.lib().flou_b200_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float(0)
        L.check(L.lib().flou_b200_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def comm_init(self, unique_id: bytes):
        L.check(L.lib().flou_b200_comm_init(self.handle, unique_id))

    def close(self):
        if getattr(self, "_h", None):
            L.lib().flou_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _state(Q, ndofs, nv, writable=False):
    if not isinstance(Q, np.ndarray) or Q.dtype != np.float64 or not Q.flags.f_contiguous:
        if writable:
            raise ValueError("output state must be a float64 column-major (ndofs, nv) array")
        Q = np.asfortranarray(Q, dtype=np.float64)
    if Q.shape != (ndofs, nv):
        raise ValueError(f"state must have shape ({ndofs}, {nv}), got {Q.shape}")   # DimensionMismatch
    return Q


def rhs(dQ, Q, p, time=0.0):
    """rhs!(dQ, Q, p::EquationConfig, time): host in, host out (parity path)."""
    disc = p.disc
    Q = _state(Q, disc.ndofs, disc.nv)
    dQ = _state(dQ, disc.ndofs, disc.nv, writable=True)
    if disc.has_dynamic:
        # source / boundary closures evaluated by the host for this (Q, time), then the device pass
        disc.upload(Q)
        disc.refresh(time, Q)
        L.check(L.lib().flou_b200_rhs(disc.handle, None, _ptr(dQ), float(time)))
        return None
    L.check(L.lib().flou_b200_rhs(disc.handle, _ptr(Q), _ptr(dQ), float(time)))
    return None


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    L.check(L.lib().flou_b200_nccl_unique_id(buf))
    return buf.raw
