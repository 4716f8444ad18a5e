"""TEST INFRASTRUCTURE ONLY — structure readers, synthetic boxes and a CPU neighbor list.

The reference obtains its neighbor list from the unvendored C library
``vesin >=0.6.1,<0.7`` (call site ``src/metatrain/utils/neighbor_lists.py:131``,
``vesin.ase_neighbor_list("ijSD", atoms, cutoff)``); its own tests only check label
names/ranks there (``tests/utils/test_neighbor_list.py:12-40``) => **parity unpinned**
for neighbor-list values.  What PET consumes is, by definition, the *set* of ordered
pairs ``(i, j, S)`` with ``|r_j + S.cell - r_i| <= cutoff`` excluding ``(i, i, 0)``
(``pet/modules/structures.py:53,71-73`` only read the samples), and that set is pinned
downstream by the energy goldens.  This file restates that definition with an
independent algorithm (periodic-image replication + scipy cKDTree).
"""
import math

import numpy as np
from scipy.spatial import cKDTree

SYMBOL_TO_Z = {"H": 1, "C": 6, "N": 7, "O": 8, "F": 9, "Si": 14}


def read_xyz_frames(path, max_frames=None):
    """Plain/extended xyz reader: returns a list of dicts (Z, positions, cell, pbc)."""
    frames = []
    with open(path) as fh:
        lines = fh.read().split("\n")
    p = 0
    while p < len(lines) and lines[p].strip():
        n = int(lines[p].split()[0])
        header = lines[p + 1]
        cell = np.zeros((3, 3))
        pbc = False
        if 'Lattice="' in header:
            lat = header.split('Lattice="')[1].split('"')[0].split()
            cell = np.array([float(x) for x in lat]).reshape(3, 3)
            pbc = True
        if 'pbc="F F F"' in header:
            pbc = False
            cell = np.zeros((3, 3))
        Z, pos = [], []
        for k in range(n):
            tok = lines[p + 2 + k].split()
            Z.append(SYMBOL_TO_Z[tok[0]])
            pos.append([float(tok[1]), float(tok[2]), float(tok[3])])
        frames.append(
            dict(Z=np.array(Z, dtype=np.int64), positions=np.array(pos), cell=cell, pbc=pbc)
        )
        p += 2 + n
        if max_frames is not None and len(frames) >= max_frames:
            break
    return frames


def read_lammps_atomic(path, type_to_Z):
    """LAMMPS data file, ``atomic`` style (id type x y z), orthorhombic box."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    lo_hi = {}
    start = None
    natoms = None
    for idx, line in enumerate(lines):
        tok = line.split()
        if len(tok) == 2 and tok[1] == "atoms":
            natoms = int(tok[0])
        if len(tok) == 4 and tok[2] in ("xlo", "ylo", "zlo"):
            lo_hi[tok[2][0]] = (float(tok[0]), float(tok[1]))
        if line.startswith("Atoms"):
            start = idx + 2
    Z, pos = [], []
    for k in range(natoms):
        tok = lines[start + k].split()
        Z.append(type_to_Z[int(tok[1])])
        pos.append([float(tok[2]) - lo_hi["x"][0], float(tok[3]) - lo_hi["y"][0],
                    float(tok[4]) - lo_hi["z"][0]])
    cell = np.diag([lo_hi[a][1] - lo_hi[a][0] for a in "xyz"])
    return dict(Z=np.array(Z, dtype=np.int64), positions=np.array(pos), cell=cell, pbc=True)


def replicate(frame, reps):
    """Tile a periodic frame ``reps=(na,nb,nc)`` times along its cell vectors."""
    na, nb, nc = reps
    cell = frame["cell"]
    pos, Z = [], []
    for a in range(na):
        for b in range(nb):
            for c in range(nc):
                pos.append(frame["positions"] + a * cell[0] + b * cell[1] + c * cell[2])
                Z.append(frame["Z"])
    new_cell = cell * np.array([[na], [nb], [nc]])
    return dict(Z=np.concatenate(Z), positions=np.concatenate(pos), cell=new_cell, pbc=True)


def silicon_box(reps=2, a=5.431, sigma=0.05, seed=0):
    """Diamond-cubic Si, ``reps^3`` conventional cells, Gaussian jitter."""
    fcc = np.array([[0, 0, 0], [0, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 0.5, 0]])
    basis = np.concatenate([fcc, fcc + 0.25]) * a
    unit = dict(Z=np.full(8, 14, dtype=np.int64), positions=basis, cell=np.eye(3) * a, pbc=True)
    box = replicate(unit, (reps, reps, reps))
    rng = np.random.default_rng(seed)
    box["positions"] = box["positions"] + rng.normal(0.0, sigma, box["positions"].shape)
    return box


def neighbor_list(positions, cell, pbc, cutoff):
    """Full (both directions) neighbor list: (centers, neighbors, shifts) sorted by
    (center, neighbor, shift).  ``|r_j + S.cell - r_i| <= cutoff``, no (i,i,0)."""
    positions = np.asarray(positions, dtype=np.float64)
    n = len(positions)
    if n == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros((0, 3), dtype=np.int64)
    if not pbc:
        tree = cKDTree(positions)
        pairs = tree.query_pairs(cutoff, output_type="ndarray")
        if len(pairs) == 0:
            z = np.zeros(0, dtype=np.int64)
            return z, z, np.zeros((0, 3), dtype=np.int64)
        i = np.concatenate([pairs[:, 0], pairs[:, 1]])
        j = np.concatenate([pairs[:, 1], pairs[:, 0]])
        d = np.linalg.norm(positions[j] - positions[i], axis=1)
        keep = d <= cutoff
        i, j = i[keep], j[keep]
        S = np.zeros((len(i), 3), dtype=np.int64)
    else:
        cell = np.asarray(cell, dtype=np.float64)
        inv = np.linalg.inv(cell)
        frac = positions @ inv
        wrap = np.floor(frac).astype(np.int64)  # image that brings atom into [0,1)^3
        wrapped = (frac - wrap) @ cell
        vol = abs(np.linalg.det(cell))
        heights = [vol / np.linalg.norm(np.cross(cell[(k + 1) % 3], cell[(k + 2) % 3]))
                   for k in range(3)]
        nimg = [int(math.ceil(cutoff / h)) for h in heights]
        # keep only images that can be within `cutoff` of the home cell
        img_pos, img_idx, img_shift = [], [], []
        fw = frac - wrap
        for a in range(-nimg[0], nimg[0] + 1):
            for b in range(-nimg[1], nimg[1] + 1):
                for c in range(-nimg[2], nimg[2] + 1):
                    s = np.array([a, b, c])
                    f = fw + s
                    # distance (in units of heights) outside the unit cube
                    out = np.maximum(np.maximum(-f, f - 1.0), 0.0) * np.array(heights)
                    sel = np.nonzero((out <= cutoff + 1e-9).all(axis=1))[0]
                    if len(sel) == 0:
                        continue
                    img_pos.append(wrapped[sel] + s @ cell)
                    img_idx.append(sel)
                    img_shift.append(np.broadcast_to(s, (len(sel), 3)))
        img_pos = np.concatenate(img_pos)
        img_idx = np.concatenate(img_idx)
        img_shift = np.concatenate(img_shift)
        tree_img = cKDTree(img_pos)
        tree_home = cKDTree(wrapped)
        sp = tree_home.sparse_distance_matrix(tree_img, cutoff + 1e-6, output_type="coo_matrix")
        i = sp.row.astype(np.int64)
        k = sp.col.astype(np.int64)
        j = img_idx[k]
        # shift relative to the *unwrapped* input positions
        S = img_shift[k] - wrap[j] + wrap[i]
        r = positions[j] - positions[i] + S @ cell
        d = np.linalg.norm(r, axis=1)
        keep = (d <= cutoff) & ~((i == j) & (S == 0).all(axis=1))
        i, j, S = i[keep], j[keep], S[keep]
    order = np.lexsort((S[:, 2], S[:, 1], S[:, 0], j, i))
    return i[order], j[order], S[order].astype(np.int64)
