"""InputBounds -- mirrors ft_mpc/controllers/tools/input_bounds.py:43-76 (construct-time, CPU, like the
reference): convex hull of the 2^(#healthy) corner wrenches -> A u <= b, rows sorted by np.unique."""
import numpy as np

_CACHE: dict = {}


def hull_of_faults(D, max_thrust, faults):
    """faults: iterable of (index, intensity).  Returns (A [n_h,6], b [n_h]).  Raises scipy QhullError for
    rank-deficient fault sets (pairs (12,13), (14,15)), exactly like the reference."""
    key = tuple(sorted((int(i), float(a)) for i, a in faults))
    if key in _CACHE:
        return _CACHE[key]
    from scipy.spatial import ConvexHull
    broken = {i: a * max_thrust for i, a in key}
    # the 2^(#healthy) corner forces of :49-63 in itertools.product order, and ONE matrix-vector product per corner as
    # in the reference: a vectorised `corners @ D.T` differs in the last bit, which is enough to change the order in
    # which np.unique sorts Qhull's facet equations, i.e. the numbering of the constraints
    # (pinned by tests/golden/ref_fixtures.npz)
    import itertools
    min_max = [[broken[i], broken[i]] if i in broken else [0.0, max_thrust] for i in range(D.shape[1])]
    corners = np.array(list(itertools.product(*min_max)))
    verts = np.unique(np.array([np.matmul(D, c) for c in corners]), axis=0)                                   # :57-67
    hull = ConvexHull(verts)                                                                              # :68
    simplified = np.unique(hull.equations, axis=0)                                                        # :71
    A, b = simplified[:, :-1].copy(), -simplified[:, -1].copy()                                           # :72-73
    _CACHE[key] = (A, b)
    return A, b


class InputBounds:
    def __init__(self, model):
        faults = [(bt.index, bt.intensity) for bt in model.broken_thrusters]
        self.A, self.b = hull_of_faults(model.D, model.max_thrust, faults)

    def get_conv_hull(self):
        return self.A, self.b
