"""CPU oracle for the sparse-convolution hot path of StefOe/DPCR-AGB (MSENet14 / MSENet50).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the timed CPU arm -- never as the thing
the CUDA path routes through.

PARITY UNPINNED.  The arithmetic of this path lives in third-party dependencies that are absent
from ``/root/reference`` and cannot be installed here (no network):

* NVIDIA/MinkowskiEngine, unpinned ``git+https://github.com/NVIDIA/MinkowskiEngine`` master
  (~v0.5.4; ``README.md:100,107`` of the reference) -- SparseTensor, coordinate manager,
  convolution, pooling, broadcast;
* pyg 2.3.1 / pytorch-cluster / pytorch-scatter (``torch-points3d/env.yml:42-44``) --
  ``grid_cluster`` and ``consecutive_cluster`` behind ``GridSampling3D``.

The reference holds no test, golden vector or fixture for any of it, so this
oracle restates the *published* semantics of those libraries and anchors on the reference's own
call sites (cited per function).  The conventions chosen where upstream GPU behaviour is
race-dependent (row order of strided maps) are the deterministic CPU-MinkowskiEngine ones and are
listed in DESIGN.md.

Modules
-------
coords   numpy restatement of voxel quantisation, strided maps, kernel maps (integer, exact)
ops      torch-CPU restatement of conv / pooling / broadcast / norm arithmetic (fp32, autograd)
me_cpu   a MinkowskiEngine-shaped namespace over ``coords`` + ``ops`` so the unchanged reference
         networks (SENet.py ...) can be executed on CPU as the checker
train    loss + AdaBelief restatement for the training-step oracle
"""
