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
    # the 2^(#healthy) corner forces of :49-63 (itertools.product there), enumerated with bit patterns;
    # np.unique below sorts the vertices, so the enumeration order is immaterial
    healthy = [i for i in range(D.shape[1]) if i not in broken]
    bits = (np.arange(2 ** len(healthy))[:, None] >> np.arange(len(healthy))[None, :]) & 1
    corners = np.zeros((bits.shape[0], D.shape[1]))
    corners[:, healthy] = bits * max_thrust
    for i, f in broken.items():
        corners[:, i] = f
    verts = np.unique(corners @ D.T, axis=0)                                                              # :67
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
