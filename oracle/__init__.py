"""CPU oracle for the equilib remap hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``equilib_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and there only as the checker
or as the timed CPU baseline -- never as the product path.
"""

from .remap_oracle import (  # noqa: F401
    cube2equi,
    cube2equi_coords,
    equi2cube,
    equi2cube_coords,
    equi2equi,
    equi2equi_coords,
    equi2pers,
    equi2pers_coords,
    pers2equi,
    pers2equi_coords,
    rotation_matrices,
    sample,
)
