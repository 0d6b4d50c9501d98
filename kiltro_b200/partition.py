"""Row-block (mesh-strip) partition of a structured node grid across ranks (new capability: the reference is
single-process).  Node numbering is x-fastest row-major (kiltro/MeshGeneration/Mesh.py:37-41), so contiguous blocks of DOFs
are horizontal strips of mesh rows."""


class StripPartition(object):
    """Mesh rows [j0, j1) of an (nx_nodes x ny_nodes) node grid owned by `rank`, plus one halo row per interior side."""

    def __init__(self, nx_nodes, ny_nodes, world, rank):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank/world")
        if ny_nodes < 2 * world:
            raise ValueError("need at least two mesh rows per rank")
        self.nxn, self.nyn, self.world, self.rank = int(nx_nodes), int(ny_nodes), int(world), int(rank)
        base, rem = divmod(self.nyn, self.world)
        self.j0 = rank * base + min(rank, rem)
        self.j1 = self.j0 + base + (1 if rank < rem else 0)
        self.hb = 1 if rank > 0 else 0                  # halo row below (owned by rank-1)
        self.ha = 1 if rank < world - 1 else 0          # halo row above (owned by rank+1)
        self.j_first = self.j0 - self.hb
        self.nrows_loc = self.j1 - self.j0 + self.hb + self.ha
        self.n_own = (self.j1 - self.j0) * self.nxn
        self.n_loc = self.nrows_loc * self.nxn
        self.own_off = self.hb * self.nxn
        self.global_off = self.j0 * self.nxn            # global node id of the first owned node

    # slices of a full local vector
    def own(self):
        return slice(self.own_off, self.own_off + self.n_own)

    def first_own_row(self):
        return slice(self.own_off, self.own_off + self.nxn)

    def last_own_row(self):
        return slice(self.own_off + self.n_own - self.nxn, self.own_off + self.n_own)

    def halo_below(self):
        return slice(0, self.nxn) if self.hb else None

    def halo_above(self):
        return slice(self.own_off + self.n_own, self.own_off + self.n_own + self.nxn) if self.ha else None
