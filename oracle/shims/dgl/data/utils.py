"""Test-infrastructure shim (NOT product code)."""


def save_info(path, info):
    """Companion of dgl.save_graphs (graph_generator.py:896-899): a rebuild cache, nothing is written (see save_graphs)."""
    return None


def load_info(path):
    raise NotImplementedError("dgl shim: no graph cache is ever written, so none can be loaded")
