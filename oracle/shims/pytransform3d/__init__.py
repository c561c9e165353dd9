"""Test-infrastructure shim (NOT product code): minimal stand-in for pytransform3d 1.9.1
so that the reference's tm_*.pickle files unpickle in a container without the package."""
