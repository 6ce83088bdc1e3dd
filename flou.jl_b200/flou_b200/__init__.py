"""flou_b200 -- Python host mirror of Flou.jl's API over the B200 CUDA library.

Same vocabulary as the reference (FlouCommon / FlouSpatial / FlouTime); every numerical
operation is executed by libflou_b200.so (hand-written sm_100a kernels) through the C ABI
in include/flou_b200.h.  No CPU fallback exists.
"""
from ._lib import DomainError, FlouB200Error, LIB_PATH, device_count, lib
from .disc import EquationConfig, MultielementDisc, nccl_unique_id, rhs
from .equations import (ChandrasekharAverage, EulerEquation, EulerInflowBC, EulerOutflowBC,
                        EulerSlipBC, Frame, GenericBC, LinearAdvection, Source, LxF, MatrixDissipation,
                        HybridDivOperator, ScalarDissipation, SplitDivOperator, StdAverage, StrongDivOperator,
                        gaussian_bump, normal_shockwave, nvariables, soundvelocity, spatialdim,
                        vars_prim2cons)
from .io import (FlouFile, add_celldata, add_fielddata, add_pointdata, add_solution, close_file,
                 get_save_callback, open_for_write, pointdata2VTKHDF, vtk_connectivities, vtk_type)
from .gmshmesh import RawHexMesh, RawMesh, UnstructuredMesh, read_msh, refine, write_msh
from .monitors import (MonitorOutput, get_cfl_callback, get_limiter, get_limiter_callback, get_monitor,
                       get_monitor_callback, list_limiters, list_monitors, make_callback_list)
from .mesh import CartesianMesh, apply_periodicBCs, partition_offsets
from .stdregions import DGSEMrec, LagrangeBasis, StdHex, StdQuad, StdSegment
from .time import CarpenterKennedy2N54, ORK256, Solution, advance, get_max_dt, timeintegrate
