"""Host-side mirror of the plugin's operator interface for the 2B+3B path.

Names, argument meaning and error behaviour follow the reference:

  MBPolTwoBodyForce / MBPolThreeBodyForce    openmmapi/include/openmm/MBPolTwoBodyForce.h:57-166, MBPolThreeBodyForce.h:57-166
                                             (SWIG classes python/mbpolplugin.i:120-171): addParticle, getNumMolecules,
                                             getParticleParameters, setParticleParameters, setCutoff, getCutoff,
                                             setNonbondedMethod, getNonbondedMethod, updateParametersInContext
  CalcMBPolTwoBodyForceKernel / ...Three...  openmmapi/include/openmm/mbpolKernels.h:134-207: Name(), initialize, execute,
                                             copyParametersToContext  -- implemented by B200CalcMBPol*ForceKernel below,
                                             the classes a CUDA-platform KernelFactory would hand out
                                             (platforms/cuda/src/CudaMBPolKernelFactory.cpp:71-82)
  MBPolContext                               the slice of OpenMM's Context the path touches: positions, periodic box,
                                             forces and energy (State::Forces | State::Energy)

All compute goes through the C ABI (include/mbpol_b200.h) in libmbpol_b200.so; positions and forces are torch CUDA
tensors or host numpy arrays.  OpenMM itself is not installed in this environment; the C++ adapter that
plugs the same C ABI into a real OpenMM is described in INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MBPolError


class _MBPolManyBodyForce:
    """Parameter container shared by the two forces (MBPolTwoBodyForce.cpp:43-81)."""

    NoCutoff = 0
    CutoffPeriodic = 1
    CutoffNonPeriodic = 2

    def __init__(self):
        # reference constructor defaults: CutoffNonPeriodic, cutoff 1e10 (MBPolTwoBodyForce.cpp:43)
        self._method = self.CutoffNonPeriodic
        self._cutoff = 1.0e10
        self._particles = []
        self._force_group = 0

    def addParticle(self, particleIndices):
        self._particles.append([int(i) for i in particleIndices])
        return len(self._particles) - 1

    def getNumParticles(self):
        return len(self._particles)

    def getNumMolecules(self):
        return len(self._particles)

    def getParticleParameters(self, particleIndex):
        return list(self._particles[particleIndex])

    def setParticleParameters(self, particleIndex, particleIndices):
        self._particles[particleIndex] = [int(i) for i in particleIndices]

    def setCutoff(self, cutoff):
        self._cutoff = float(cutoff)

    def getCutoff(self):
        return self._cutoff

    def getNonbondedMethod(self):
        return self._method

    def setNonbondedMethod(self, method):
        if method not in (0, 1, 2):
            raise ValueError("unknown NonbondedMethod")
        self._method = int(method)

    def getForceGroup(self):
        return self._force_group

    def setForceGroup(self, group):
        self._force_group = int(group)

    def updateParametersInContext(self, context):
        context._update_parameters(self)

    def _indices(self):
        if any(len(p) < 3 for p in self._particles):
            raise MBPolError(_lib.EINVAL, "each molecule needs the particle indices {O, H1, H2}")
        return np.asarray([p[:3] for p in self._particles], dtype=np.int32).reshape(-1, 3)


class MBPolTwoBodyForce(_MBPolManyBodyForce):
    """openmmapi/include/openmm/MBPolTwoBodyForce.h:57-166"""


class MBPolThreeBodyForce(_MBPolManyBodyForce):
    """openmmapi/include/openmm/MBPolThreeBodyForce.h:57-166"""


class MBPolOneBodyForce:
    """MBPolOneBodyForce (openmmapi/include/openmm/MBPolOneBodyForce.h; SWIG python/mbpolplugin.i): the intramolecular
    Partridge-Schwenke term, one {O,H1,H2} entry per molecule."""

    NonPeriodic = 0
    Periodic = 1

    def __init__(self):
        self._method = self.NonPeriodic
        self._bodies = []
        self._group = 0

    def addOneBody(self, particleIndices):
        p = [int(i) for i in particleIndices]
        if len(p) != 3:
            raise MBPolError(_lib.EINVAL, "addOneBody: three particle indices {O,H1,H2} are required")
        self._bodies.append(p)
        return len(self._bodies) - 1

    def getNumOneBodys(self):
        return len(self._bodies)

    def getOneBodyParameters(self, index):
        return list(self._bodies[index])

    def setOneBodyParameters(self, index, particleIndices):
        self._bodies[index] = [int(i) for i in particleIndices]

    def getNonbondedMethod(self):
        return self._method

    def setNonbondedMethod(self, method):
        if method not in (self.NonPeriodic, self.Periodic):
            raise MBPolError(_lib.EINVAL, "MBPolOneBodyForce: Illegal value for nonbonded method")
        self._method = method

    def getForceGroup(self):
        return self._group

    def setForceGroup(self, group):
        self._group = int(group)

    def getNumParticles(self):
        return len(self._bodies)

    def updateParametersInContext(self, context):
        for k in context.kernels:
            if k.force is self:
                k.copyParametersToContext(context, self)

    def _indices(self):
        return np.asarray(self._bodies, dtype=np.int32).reshape(-1, 3)


class CustomDispersionForce:
    """The OpenMM CustomNonbondedForce that python/mbpol.xml:37-108 builds for MB-pol's TT-damped C6 dispersion
    (-C6 tt6(r) / r^6 between O / H atoms of different molecules, M sites excluded; tables mbpol.xml:53-82), as a
    parameter object: nonbonded method, cutoff distance (on the atom-atom distance) and one {O, H1, H2} entry per water."""

    NoCutoff = 0
    CutoffNonPeriodic = 1
    CutoffPeriodic = 2

    def __init__(self):
        self._method = self.NoCutoff
        self._cutoff = 1.0                          # CustomNonbondedForce default cutoff distance, nm
        self._molecules = []
        self._group = 0

    def addMolecule(self, particleIndices):
        p = [int(i) for i in particleIndices]
        if len(p) < 3:
            raise MBPolError(_lib.EINVAL, "addMolecule: the particle indices {O, H1, H2} are required")
        self._molecules.append(p[:3])
        return len(self._molecules) - 1

    def getNumMolecules(self):
        return len(self._molecules)

    def getNumParticles(self):
        return len(self._molecules)

    def getNonbondedMethod(self):
        return self._method

    def setNonbondedMethod(self, method):
        if method not in (0, 1, 2):
            raise MBPolError(_lib.EINVAL, "CustomNonbondedForce: Illegal value for nonbonded method")
        self._method = int(method)

    def getCutoffDistance(self):
        return self._cutoff

    def setCutoffDistance(self, distance):
        self._cutoff = float(distance)

    def getForceGroup(self):
        return self._group

    def setForceGroup(self, group):
        self._group = int(group)

    def _indices(self):
        return np.asarray(self._molecules, dtype=np.int32).reshape(-1, 3)


DO_1B = 0x100      # host-side selectors only: these terms have their own C-ABI entries (mbpol_eval_onebody,
DO_DISP = 0x200    # mbpol_eval_dispersion)


class _Engine:
    """One mbpol_ctx shared by the 2B and 3B kernels of a context (one list build per step serves both)."""

    def __init__(self, n_particles, idx, cut2, cut3, m2, m3, device=-1):
        self.L = _lib.load()
        self.idx = np.ascontiguousarray(idx, dtype=np.int32)
        cfg = _lib.Config(_lib.ABI_VERSION, len(self.idx), int(n_particles),
                          self.idx.ctypes.data_as(C.POINTER(C.c_int32)), float(cut2), float(cut3), int(m2), int(m3),
                          int(device), 0)
        h = C.c_void_p()
        _lib.check(self.L.mbpol_create(C.byref(h), C.byref(cfg)))
        self.h = h
        self.n_particles = int(n_particles)

    def close(self):
        if getattr(self, "h", None):
            self.L.mbpol_destroy(self.h)
            self.h = None

    __del__ = close

    def set_box(self, box):
        b = (C.c_double * 3)(*[float(x) for x in box])
        _lib.check(self.L.mbpol_set_box(self.h, b), self.h)

    def set_skin(self, skin_nm):
        """Verlet skin for the neighbour rows (0 = rebuild every evaluation, the reference's behaviour)"""
        _lib.check(self.L.mbpol_set_skin(self.h, float(skin_nm)), self.h)

    def set_partition(self, rank, world):
        _lib.check(self.L.mbpol_set_partition(self.h, rank, world), self.h)

    def update_particles(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        _lib.check(self.L.mbpol_update_particles(self.h, len(idx), idx.ctypes.data_as(C.POINTER(C.c_int32))), self.h)

    def eval_device(self, pos, force, flags, want_energy=True, stream=None):
        """pos/force: torch float64 CUDA tensors (n_particles x 3); force is accumulated into."""
        e = (C.c_double * 2)()
        _lib.check(self.L.mbpol_eval(self.h, pos.data_ptr(), force.data_ptr() if force is not None else None,
                                     e if want_energy else None, flags, stream), self.h)
        return (e[0], e[1]) if want_energy else None

    def eval_host(self, pos, force, flags):
        """pos/force: C-contiguous float64 numpy arrays (n_particles x 3); force is accumulated into."""
        e = (C.c_double * 2)()
        _lib.check(self.L.mbpol_eval_host(self.h, pos.ctypes.data, force.ctypes.data if force is not None else None,
                                          e, flags), self.h)
        return e[0], e[1]

    def eval_onebody(self, pos, force, periodic, stream=None):
        """one-body term; pos/force torch CUDA tensors or host numpy arrays, force accumulated into"""
        e = C.c_double()
        if isinstance(pos, np.ndarray):
            _lib.check(self.L.mbpol_eval_onebody_host(self.h, pos.ctypes.data,
                                                      force.ctypes.data if force is not None else None,
                                                      C.byref(e), int(periodic)), self.h)
        else:
            _lib.check(self.L.mbpol_eval_onebody(self.h, pos.data_ptr(),
                                                 force.data_ptr() if force is not None else None, C.byref(e),
                                                 int(periodic), stream), self.h)
        return e.value

    def eval_dispersion(self, pos, force, cutoff, periodic, stream=None):
        """TT-damped C6 dispersion; pos/force torch CUDA tensors or host numpy arrays, force accumulated into"""
        e = C.c_double()
        if isinstance(pos, np.ndarray):
            _lib.check(self.L.mbpol_eval_dispersion_host(self.h, pos.ctypes.data,
                                                         force.ctypes.data if force is not None else None, C.byref(e),
                                                         float(cutoff), int(periodic)), self.h)
        else:
            _lib.check(self.L.mbpol_eval_dispersion(self.h, pos.data_ptr(), force.data_ptr() if force is not None else None,
                                                    C.byref(e), float(cutoff), int(periodic), stream), self.h)
        return e.value

    def _get_list(self, fn, width):
        n = C.c_size_t(0)
        _lib.check(fn(self.h, None, C.byref(n)), self.h)
        out = np.empty((n.value, width), dtype=np.int32)
        _lib.check(fn(self.h, out.ctypes.data, C.byref(n)), self.h)
        return out

    def pairs(self):
        return self._get_list(self.L.mbpol_get_pairs, 2)

    def triplets(self):
        return self._get_list(self.L.mbpol_get_triplets, 3)

    def stats(self):
        s = _lib.Stats()
        _lib.check(self.L.mbpol_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _lib.Stats._fields_}

    def set_timing(self, on):
        _lib.check(self.L.mbpol_set_timing(self.h, int(bool(on))))


class _B200Kernel:
    """Common part of the two platform kernels (what CudaCalcMBPol*ForceKernel is in the reference's CUDA tree)."""

    _flag = 0

    def __init__(self, name, context):
        self.name = name
        self.context = context
        self.force = None

    def initialize(self, system, force):
        # ReferenceCalcMBPol*ForceKernel::initialize, MBPolReferenceKernels.cpp:341-360 / 430-449
        self.force = force
        self.context._register(self, force)

    def execute(self, context, includeForces=True, includeEnergy=True):
        # like the reference, forces and energy are always both computed (MBPolReferenceKernels.cpp:362-399)
        return context._execute(self._flag)

    def copyParametersToContext(self, context, force):
        if force.getNumParticles() != self.force.getNumParticles():
            raise MBPolError(_lib.ECHANGED, "updateParametersInContext: The number of particles has changed")
        context._update_parameters(force)


class B200CalcMBPolTwoBodyForceKernel(_B200Kernel):
    _flag = _lib.DO_2B

    @staticmethod
    def Name():
        return "CalcMBPolTwoBodyForce"


class B200CalcMBPolThreeBodyForceKernel(_B200Kernel):
    _flag = _lib.DO_3B

    @staticmethod
    def Name():
        return "CalcMBPolThreeBodyForce"


class B200CalcMBPolOneBodyForceKernel(_B200Kernel):
    """CalcMBPolOneBodyForceKernel (openmmapi/include/openmm/mbpolKernels.h:47-82; Reference implementation
    MBPolReferenceKernels.cpp:77-126)"""
    _flag = DO_1B

    @staticmethod
    def Name():
        return "CalcMBPolOneBodyForce"

    def copyParametersToContext(self, context, force):
        if force.getNumOneBodys() != self.force.getNumOneBodys():
            raise MBPolError(_lib.ECHANGED, "updateParametersInContext: The number of stretch-bends has changed")
        context._update_parameters(force)


class B200CalcCustomDispersionForceKernel(_B200Kernel):
    """stands where OpenMM's CalcCustomNonbondedForceKernel would evaluate the dispersion force of mbpol.xml:37-108"""
    _flag = DO_DISP

    @staticmethod
    def Name():
        return "CustomDispersionForce"

    def copyParametersToContext(self, context, force):
        if force.getNumMolecules() != self.force.getNumMolecules():
            raise MBPolError(_lib.ECHANGED, "updateParametersInContext: The number of particles has changed")
        context._update_parameters(force)


class MBPolContext:
    """The part of OpenMM's System+Context the 2B/3B path needs.

        ctx = MBPolContext(n_particles, [two_body_force, three_body_force], box_nm=(L, L, L))
        ctx.setPositions(pos_nm)                       # numpy (host) or torch CUDA tensor, n_particles x 3
        energy = ctx.getPotentialEnergy()              # kJ/mol  (dict per force via ctx.energies)
        forces = ctx.getForces()                       # kJ/mol/nm, same kind of array as the positions
    """

    def __init__(self, n_particles, forces, box_nm=None, device=-1, rank=0, world=1):
        self.n_particles = int(n_particles)
        self.forces = list(forces)
        self.box = None if box_nm is None else [float(x) for x in box_nm]
        self._f2 = next((f for f in self.forces if isinstance(f, MBPolTwoBodyForce)), None)
        self._f3 = next((f for f in self.forces if isinstance(f, MBPolThreeBodyForce)), None)
        self._f1 = next((f for f in self.forces if isinstance(f, MBPolOneBodyForce)), None)
        self._fd = next((f for f in self.forces if isinstance(f, CustomDispersionForce)), None)
        if self._f2 is None and self._f3 is None and self._f1 is None and self._fd is None:
            raise MBPolError(_lib.EINVAL, "no MBPolOneBodyForce / MBPolTwoBodyForce / MBPolThreeBodyForce given")
        ref = self._f2 or self._f3 or self._f1 or self._fd
        if self._fd is not None:
            if not np.array_equal(self._fd._indices(), ref._indices()):
                raise MBPolError(_lib.EINVAL, "the forces must list the same molecules")
            if self._fd.getNonbondedMethod() == CustomDispersionForce.CutoffPeriodic and self.box is None:
                raise MBPolError(_lib.EINVAL, "CutoffPeriodic needs a periodic box")
        if self._f1 is not None and not np.array_equal(self._f1._indices(), ref._indices()):
            raise MBPolError(_lib.EINVAL, "the forces must list the same molecules")
        if self._f1 is not None and self._f1.getNonbondedMethod() == MBPolOneBodyForce.Periodic and self.box is None:
            raise MBPolError(_lib.EINVAL, "MBPolOneBodyForce.Periodic needs a periodic box")
        if self._f2 is not None and self._f3 is not None and not np.array_equal(self._f2._indices(), self._f3._indices()):
            raise MBPolError(_lib.EINVAL, "the two forces must list the same molecules")
        # MBPol*ForceImpl::initialize: cutoff <= box/2 for periodic systems (MBPolTwoBodyForceImpl.cpp:60-66)
        for f in (self._f2, self._f3):
            if f is not None and f.getNonbondedMethod() == f.CutoffPeriodic:
                if self.box is None:
                    raise MBPolError(_lib.EINVAL, "CutoffPeriodic needs a periodic box")
                if f.getCutoff() > 0.5 * min(self.box):
                    raise MBPolError(_lib.EBOX, "MBPolForce: The cutoff distance cannot be greater than half the "
                                                "periodic box size.")
        d = lambda f, attr, default: getattr(f, attr)() if f is not None else default  # noqa: E731
        self.engine = _Engine(self.n_particles, ref._indices(), d(self._f2, "getCutoff", 1e10),
                              d(self._f3, "getCutoff", 1e10), d(self._f2, "getNonbondedMethod", 2),
                              d(self._f3, "getNonbondedMethod", 2), device)
        if self.box is not None:
            self.engine.set_box(self.box)
        if world > 1:
            self.engine.set_partition(rank, world)
        self.kernels = []
        self._flags = 0
        for f, K in ((self._f1, B200CalcMBPolOneBodyForceKernel), (self._f2, B200CalcMBPolTwoBodyForceKernel),
                     (self._f3, B200CalcMBPolThreeBodyForceKernel), (self._fd, B200CalcCustomDispersionForceKernel)):
            if f is not None:
                k = K(K.Name(), self)
                k.initialize(None, f)
        self._pos = None
        self.energies = {}
        self._forces = None
        self._force_out = None

    # -- plumbing used by the kernels ---------------------------------------------------------------------
    def _register(self, kernel, force):
        self.kernels.append(kernel)
        self._flags |= kernel._flag

    def _update_parameters(self, force):
        # The forces of one context share one engine, hence one molecule table.  An update of one force is pushed at
        # once; that all forces list the same molecules again (what the constructor enforces) is checked at the next
        # evaluation, so that the forces can be updated one after the other.
        self.engine.update_particles(force._indices())
        self._check_same_molecules = True

    def _same_molecules_or_raise(self):
        ref = self._f2 or self._f3 or self._f1 or self._fd
        for other in (self._f1, self._f2, self._f3, self._fd):
            if other is not None and other is not ref and not np.array_equal(other._indices(), ref._indices()):
                raise MBPolError(_lib.EINVAL, "updateParametersInContext: the forces of this context no longer list "
                                              "the same molecules")
        self.engine.update_particles(ref._indices())
        self._check_same_molecules = False

    def _execute(self, flags):
        if self._pos is None:
            raise MBPolError(_lib.EINVAL, "positions have not been set")
        if getattr(self, "_check_same_molecules", False):
            self._same_molecules_or_raise()
        host = isinstance(self._pos, np.ndarray)
        stream = None
        many = flags & (_lib.DO_2B | _lib.DO_3B)
        overwrite = host and many and not flags & (DO_1B | DO_DISP)   # the many-body call may then write f without reading it
        if host:
            if self._force_out is not None:
                f = self._force_out
                if not overwrite:
                    f[...] = 0.0
            else:
                f = np.empty_like(self._pos) if overwrite else np.zeros_like(self._pos)
        else:
            import torch
            f = torch.zeros_like(self._pos)
            stream = torch.cuda.current_stream().cuda_stream
        self.energies = {}
        total = 0.0
        if flags & DO_1B:
            e1 = self.engine.eval_onebody(self._pos, f, self._f1.getNonbondedMethod() == MBPolOneBodyForce.Periodic,
                                          stream)
            self.energies["CalcMBPolOneBodyForce"] = e1
            total += e1
        if flags & DO_DISP:
            d = self._fd
            cutoff = 1e10 if d.getNonbondedMethod() == d.NoCutoff else d.getCutoffDistance()
            ed = self.engine.eval_dispersion(self._pos, f, cutoff, d.getNonbondedMethod() == d.CutoffPeriodic, stream)
            self.energies["CustomDispersionForce"] = ed
            total += ed
        if many:
            e2, e3 = self.engine.eval_host(self._pos, f, many | (_lib.OVERWRITE_FORCES if overwrite else 0)) if host else \
                self.engine.eval_device(self._pos, f, many, stream=stream)
            self.energies.update({"CalcMBPolTwoBodyForce": e2, "CalcMBPolThreeBodyForce": e3})
            total += e2 + e3
        self._forces = f
        return total

    # -- user-facing ----------------------------------------------------------------------------------------
    def setPeriodicBoxVectors(self, a, b=None, c=None):
        box = [a[0], b[1], c[2]] if b is not None else list(a)
        self.box = [float(x) for x in box]
        self.engine.set_box(self.box)

    def setPositions(self, positions):
        if isinstance(positions, np.ndarray) or not hasattr(positions, "data_ptr"):
            p = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        else:
            p = positions.contiguous()
            if p.dtype is not __import__("torch").float64 or not p.is_cuda:
                raise MBPolError(_lib.EINVAL, "device positions must be a float64 CUDA tensor")
            p = p.reshape(-1, 3)
        if p.shape[0] != self.n_particles:
            raise MBPolError(_lib.EINVAL, "Called setPositions() with the wrong number of positions")
        self._pos = p
        self._forces = None

    def setForceOutput(self, array):
        """Host evaluations write their forces into this (e.g. pinned) n_particles x 3 float64 array instead of a new one."""
        if array is not None and (array.shape != (self.n_particles, 3) or array.dtype != np.float64 or
                                  not array.flags.c_contiguous):
            raise MBPolError(_lib.EINVAL, "force output must be a C-contiguous float64 array of shape (n_particles, 3)")
        self._force_out = array

    def calcForcesAndEnergy(self, groups=0xFFFFFFFF):
        """ContextImpl::calcForcesAndEnergy for the registered kernels, honouring force groups
        (MBPolTwoBodyForceImpl.cpp:72-76).  One C-ABI call evaluates all selected terms on one tuple-list build."""
        flags = 0
        for k in self.kernels:
            if groups & (1 << k.force.getForceGroup()):
                flags |= k._flag
        if not flags:
            self.energies = {}
            return 0.0
        return self._execute(flags)

    def getPotentialEnergy(self, groups=0xFFFFFFFF):
        return self.calcForcesAndEnergy(groups)

    def getForces(self):
        if self._forces is None:
            self.calcForcesAndEnergy()
        return self._forces

    def getState(self, groups=0xFFFFFFFF):
        e = self.calcForcesAndEnergy(groups)
        return {"energy": e, "forces": self._forces, **self.energies}

    def getPairs(self):
        return self.engine.pairs()

    def getTriplets(self):
        return self.engine.triplets()


def pip_eval(order, x):
    """Bare PIP energy + gradient for a batch of variable vectors: x = torch float64 CUDA tensor [n_vars, n]."""
    import torch
    L = _lib.load()
    nv = {2: 31, 3: 36}[order]
    assert x.shape[0] == nv and x.is_cuda and x.dtype == torch.float64
    x = x.contiguous()
    out = torch.empty(nv + 1, x.shape[1], dtype=torch.float64, device=x.device)
    _lib.check(L.mbpol_pip_eval(order, x.data_ptr(), out.data_ptr(), x.shape[1], x.shape[1],
                                torch.cuda.current_stream().cuda_stream))
    return out[0], out[1:]


def eval_tuples(order, pos_mol, tuples, box=None):
    """Per-tuple energies and force blocks for explicit tuples over stride-3 molecules (pos_mol: [n_mol,3,3] nm)."""
    import torch
    L = _lib.load()
    pos = torch.as_tensor(np.ascontiguousarray(pos_mol, dtype=np.float64)).cuda().reshape(-1, 3)
    tup = torch.as_tensor(np.ascontiguousarray(tuples, dtype=np.int32)).cuda()
    n = tup.shape[0]
    e = torch.zeros(n, dtype=torch.float64, device="cuda")
    f = torch.zeros(n, 3 * order, 3, dtype=torch.float64, device="cuda")
    b = (C.c_double * 3)(*(box if box is not None else (0.0, 0.0, 0.0)))
    _lib.check(L.mbpol_eval_tuples(order, pos.data_ptr(), pos.shape[0] // 3, tup.data_ptr(), n, b,
                                   int(box is not None), e.data_ptr(), f.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))
    return e.cpu().numpy(), f.cpu().numpy()


def measure_fp64_peak():
    L = _lib.load()
    t, ms = C.c_double(), C.c_double()
    _lib.check(L.mbpol_measure_fp64_peak(C.byref(t), C.byref(ms)))
    return t.value, ms.value
