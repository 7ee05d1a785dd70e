"""InputBounds -- mirrors ft_mpc/controllers/tools/input_bounds.py:43-76 (construct-time, CPU, like the
reference): convex hull of the 2^(#healthy) corner wrenches -> A u <= b, rows sorted by np.unique."""
import numpy as np

_CACHE: dict = {}


def hull_of_faults(D, max_thrust, faults, method="qhull"):
    """faults: iterable of (index, intensity).  Returns (A [n_h,6], b [n_h]).
    method "qhull" (default) follows the reference line by line, including the row ORDER np.unique gives Qhull's
    equations (it is driven by Qhull's rounding noise, so it is the only way to number the constraints like the
    reference does) and raises scipy's QhullError for rank-deficient fault sets (pairs (12,13), (14,15)).
    method "analytic" enumerates the zonotope's facets directly (`zonotope_facets`: same facet set, canonical row order,
    ~15 ms instead of ~1 s, any intensity, rank-deficient sets included)."""
    if method == "analytic":
        A, b, _ = zonotope_facets(D, max_thrust, faults)
        return A, b
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


# ---------------------------------------------------------------------------------------------------------
# Analytic construction: the wrench set is a ZONOTOPE
#     { D f_fault + sum_{i healthy} u_i D[:, i],  0 <= u_i <= max_thrust },
# so its facets need no 2^14-point convex hull: every facet normal is orthogonal to d - 1 independent generators.
# All (d-1)-subsets of the healthy generators are enumerated (C(15,5) = 3003 null vectors as batched cofactor
# expansions), duplicate hyperplanes are merged, and the offsets come from the support function
#     h(n) = n . D f_fault + max_thrust * sum_i max(0, n . D[:, i]).
# A depends only on the fault MASK; intensities only shift b -- continuous intensities need no table.
# Rank-deficient cells (pairs (12,13), (14,15): Qhull raises) are handled in the span of the generators and return
# the flat directions as pairs of opposite rows (zero width).
# ---------------------------------------------------------------------------------------------------------
def _null_vectors(M):
    """M [K, r-1, r] -> [K, r] generalised cross products (cofactors); zero vector when the rows are dependent."""
    K, _, r = M.shape
    out = np.empty((K, r))
    cols = np.arange(r)
    for j in range(r):
        out[:, j] = (-1.0) ** j * np.linalg.det(M[:, :, cols != j])
    return out


def zonotope_facets(D, max_thrust, faults, tol=1e-9):
    """Facets of the wrench zonotope of a fault set.  Returns (A [n_h,6], b [n_h], rank): unit normals, rows ordered
    like np.unique orders Qhull's equations [A | -b] (lexicographic; values compared after rounding to `tol`)."""
    import itertools
    D = np.asarray(D, dtype=float)
    broken = {int(i): float(a) * max_thrust for i, a in faults}
    healthy = [i for i in range(D.shape[1]) if i not in broken]
    offset = sum((D[:, i] * f for i, f in broken.items()), np.zeros(D.shape[0]))
    G = D[:, healthy]                                             # generators (columns)
    # orthonormal basis of the span (rank r); flat directions complete it
    U, sv, _ = np.linalg.svd(G, full_matrices=True)
    r = int((sv > 1e-9 * sv[0]).sum())
    Q, Qperp = U[:, :r], U[:, r:]
    Gr = Q.T @ G                                                  # [r, m]
    rows = []
    if r >= 2:
        subs = np.array(list(itertools.combinations(range(len(healthy)), r - 1)))
        nv = _null_vectors(np.transpose(Gr[:, subs], (1, 2, 0)))  # [K, r-1, r] -> [K, r]
        norm = np.linalg.norm(nv, axis=1)
        scale = np.abs(np.linalg.norm(Gr, axis=0)).max() ** (r - 1)
        nv = nv[norm > 1e-9 * scale] / norm[norm > 1e-9 * scale, None]
        # one representative per hyperplane: first significant component positive
        sign = np.sign(nv[np.arange(len(nv)), np.argmax(np.abs(nv) > 1e-6, axis=1)])
        nv = nv * sign[:, None]
        _, first = np.unique(np.round(nv / tol) * tol, axis=0, return_index=True)
        for n_r in nv[np.sort(first)]:
            n = Q @ n_r
            for sgn in (1.0, -1.0):
                nn = sgn * n
                rows.append(np.append(nn, nn @ offset + max_thrust * np.maximum(0.0, nn @ G).sum()))
    elif r == 1:
        for sgn in (1.0, -1.0):
            nn = sgn * Q[:, 0]
            rows.append(np.append(nn, nn @ offset + max_thrust * np.maximum(0.0, nn @ G).sum()))
    for k in range(Qperp.shape[1]):                               # flat directions: n . y = n . offset
        for sgn in (1.0, -1.0):
            nn = sgn * Qperp[:, k]
            rows.append(np.append(nn, nn @ offset))
    rows = np.array(rows)
    A, b = rows[:, :-1], rows[:, -1]
    A = np.where(np.abs(A) < tol, 0.0, A)
    key = np.round(np.column_stack([A, -b]) / tol) * tol + 0.0    # "+ 0.0": -0.0 -> 0.0
    order = np.lexsort(key.T[::-1])
    return A[order], b[order], r
