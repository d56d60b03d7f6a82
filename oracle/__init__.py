"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the POY5 dynamic-homology alignment hot path.  Nothing under
``poy5_b200/`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do.

* ``oracle.refbind``            -- ctypes binding of ``oracle/_ref/libpoyref*.so`` (the
  UNMODIFIED reference C of amnh/poy5 compiled from ``/root/reference/src``).
* ``oracle.port``               -- ctypes binding of ``oracle/libdooracle.so`` (plain-C
  restatement ``oracle/do_oracle.c``; travels to the GPU box as source + .so).
* ``oracle.cost_matrix_oracle`` -- numpy/python restatement of ``src/cost_matrix.ml``
  table construction (OCaml cannot run here).
"""
