"""TEST INFRASTRUCTURE ONLY.  Loads the reference's own, unmodified Davidson-Liu solver
(/root/reference/forte2/helpers/davidsonliu.py) by path, without importing the `forte2` package
(which needs the compiled forte2.lib): four stub modules satisfy the imports of
helpers/logger.py, helpers/table.py and base_classes/params.py.  Only available in the build
container (where /root/reference exists): used to generate tests/golden/roots_*.npz and by CPU tests.
"""
import importlib.util
import os
import sys
import types

REF = "/root/reference/forte2"


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "helpers", "davidsonliu.py"))


def load():
    """Returns (DavidsonLiuSolver, DavidsonLiuParams) of the reference."""
    if "forte2.helpers.davidsonliu" in sys.modules:
        m = sys.modules["forte2.helpers.davidsonliu"]
        return m.DavidsonLiuSolver, sys.modules["forte2.base_classes.params"].DavidsonLiuParams

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    stub("forte2")
    stub("forte2.lib")
    stub("forte2.lib.cpp_helpers", set_log_level=lambda level: None, get_num_threads=lambda: 1)
    stub("forte2.lib.det", Determinant=type("Determinant", (), {}))
    stub("forte2.helpers")
    stub("forte2.base_classes")

    def load_file(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    logger = load_file("forte2.helpers.logger", "helpers/logger.py")
    sys.modules["forte2.helpers"].logger = logger
    load_file("forte2.helpers.table", "helpers/table.py")
    params = load_file("forte2.base_classes.params", "base_classes/params.py")
    dl = load_file("forte2.helpers.davidsonliu", "helpers/davidsonliu.py")
    logger.LoggerConfig.set_log_level(0)
    return dl.DavidsonLiuSolver, params.DavidsonLiuParams
