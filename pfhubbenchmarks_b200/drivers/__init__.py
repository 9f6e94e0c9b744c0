"""Driver scripts mirroring dolfin/bench{1,2,3,6}.py and dolfin/b13d.py of the reference."""
