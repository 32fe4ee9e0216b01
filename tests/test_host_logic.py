"""CPU tests of the closed forms the CUDA kernels rely on, against the literal loops
of the reference restated in plain Python, and of the library boundary (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np

from ndnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def literal_insert(divs):
    """kullback_leibler.c:181-195: insert before the first element strictly smaller."""
    lst = []
    for seq, d in enumerate(divs):
        j = 0
        while j < len(lst):
            if lst[j][0] < d:
                break
            j += 1
        lst.insert(j, (d, seq))
    return [s for _, s in lst]


def closed_form_order(divs):
    """What k_select does: NaN takes the exclusive prefix-min of the non-NaN values before it (+inf if none);
    stable sort by (value descending, sequence ascending); -0.0 == +0.0."""
    keys = []
    m = np.inf
    for seq, d in enumerate(divs):
        if np.isnan(d):
            keys.append((-m, seq))
        else:
            keys.append((-(d + 0.0), seq))
            m = min(m, d)
    return [s for _, s in sorted(keys)]


def test_nan_rule_closed_form_matches_literal_insertion():
    rng = np.random.default_rng(0)
    pool = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1.0, 1.0, -2.5, 3.25, 1e300, -1e300])
    for _ in range(3000):
        n = int(rng.integers(0, 40))
        divs = rng.choice(pool, n).tolist() if rng.random() < 0.7 else rng.normal(size=n).round(1).tolist()
        assert literal_insert(divs) == closed_form_order(divs)


def literal_prune(order_p, n_valid, desired):
    """ndt.c:45-67 on a list of p indices in list order."""
    alive = {}
    K = len(order_p)
    removed, idx, i = [], 0, 0
    to_remove = n_valid - desired
    ret = 0
    while i < to_remove:
        if idx >= K:
            ret = -2
            break
        p = order_p[idx]
        if alive.get(p, True):
            alive[p] = False
            removed.append(p)
            K -= 1
            i += 1
        idx += 1
    return removed, ret


def closed_form_prune(order_p, n_valid, desired):
    K = len(order_p)
    to_remove = n_valid - desired
    seen, removed, r = set(), [], 0
    for pos, p in enumerate(order_p):
        if p in seen:
            continue
        seen.add(p)
        if r < to_remove and pos + r < K:
            removed.append(p)
        r += 1
    return removed, (-2 if len(removed) < to_remove else 0)


def test_prune_walk_closed_form_matches_literal():
    rng = np.random.default_rng(1)
    for _ in range(3000):
        V = int(rng.integers(1, 30))
        K = int(rng.integers(0, 60))
        order = rng.integers(0, V, K).tolist()
        desired = int(rng.integers(0, V + 1))
        assert literal_prune(order, V, desired) == closed_form_prune(order, V, desired)


def test_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the CUDA library first (python __graft_entry__.py)"
    L = ctypes.CDLL(_lib.LIB_PATH)          # loading needs no GPU; nothing is computed here
    header = open(os.path.join(ROOT, "include", "ndnet_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b([a-z_0-9]+)\s*\(", header)) - {"defined"}
    declared = {d for d in declared if d.startswith(("ndnet_b200_", "ndt_", "prune_", "to_point", "free_", "print_"))}
    assert declared == set(_lib.EXPORTED), declared ^ set(_lib.EXPORTED)
    for name in declared:
        assert hasattr(L, name), name
    L.ndnet_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.ndnet_b200_version()


def test_product_path_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "ndt-net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "libndt_oracle" not in text, f


def test_cpulist_parser():
    from ndnet_b200.dist import _parse_cpulist
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()


def test_state_stamp_follows_in_place_updates_and_loads():
    """The eval-mode CUDA model is rebuilt when this stamp changes (ndnet/models/ndtnet.py::_state_stamp)."""
    import torch
    from ndnet.models.ndtnet import NDTNetSegmentation, _state_stamp
    net = NDTNetSegmentation(num_classes=4, feature_dim=64)
    s0 = _state_stamp(net)
    assert _state_stamp(net) == s0                                   # reading does not change it
    with torch.no_grad():
        net.conv4.bias.add_(1.0)                                     # what an optimizer step does
    s1 = _state_stamp(net)
    assert s1 != s0
    net.load_state_dict(net.state_dict())                            # copy_ into every tensor
    s2 = _state_stamp(net)
    assert s2 != s1
    net.train()
    net(torch.randn(2, 16, 3), torch.randn(2, 16, 9))                # BatchNorm running statistics move
    assert _state_stamp(net) != s2
