"""The bookkeeping of the engine's own parcel order (engine.cu do_sort / hand_over / unscramble), restated in numpy: two states
go through the same sequence of cell sorts -- one in the reference's way (module_sort permutes the parcels, what belongs to the
slot stays, src/mptrac.c:5944-5949), one with the slots sorted and the parcels laid out by another key -- and must show the
same arrays whenever they are read by slot.  The GPU test `test_the_engines_own_parcel_order_is_invisible` checks the kernels;
this one checks the index algebra, without a GPU."""
import numpy as np


class Plain:
    """position = slot"""
    def __init__(self, ident, state):
        self.ident, self.state = ident.copy(), state.copy()

    def sort(self, key_of_parcel):
        perm = np.argsort(key_of_parcel[self.ident], kind="stable")      # the stable index sort both sides use
        self.ident = self.ident[perm]                                       # parcels move, the slot-bound state stays

    def by_slot(self):
        return self.ident, self.state


class Own:
    """position != slot between two sorts: slot[i] = slot of the parcel at position i; the slot-bound state travels with the
    parcel and is handed to the slot's new parcel at a sort"""
    def __init__(self, ident, state):
        self.ident, self.state = ident.copy(), state.copy()
        self.slot = None

    def sort(self, key_of_parcel, own_key_of_parcel):
        n = self.ident.size
        held_at = np.arange(n) if self.slot is None else self._inverse(self.slot)     # invert_kernel
        key_at, own_at = key_of_parcel[self.ident], own_key_of_parcel[self.ident]     # two_keys_kernel (per position)
        order1 = np.argsort(key_at[held_at], kind="stable")                           # pick_keys + radix sort, in slot order
        takes = held_at[order1]                                                        # perm[1][s] = position of the parcel that takes slot s
        order2 = np.argsort(own_at[takes], kind="stable")                             # pick_keys + radix sort
        slot_new, src = order2, takes[order2]                                          # compose_kernel
        self.ident = self.ident[src]                                                   # gather_kernel
        self.state = self.state[held_at[slot_new]]                                     # hand_over_kernel
        self.slot = slot_new

    def unscramble(self):
        if self.slot is None:
            return
        inv = self._inverse(self.slot)
        self.ident, self.state, self.slot = self.ident[inv], self.state[inv], None

    def by_slot(self):
        if self.slot is None:
            return self.ident, self.state
        inv = self._inverse(self.slot)
        return self.ident[inv], self.state[inv]

    @staticmethod
    def _inverse(p):
        inv = np.empty_like(p)
        inv[p] = np.arange(p.size)
        return inv


def test_slots_sorted_parcels_laid_out_otherwise_is_invisible_by_slot():
    rng = np.random.default_rng(3)
    n = 5003
    ident = rng.permutation(n)
    state = rng.normal(size=n)
    a, b = Plain(ident, state), Own(ident, state)
    for event in range(12):
        key = rng.integers(0, 40, n)                # few cells: many ties, the stable order matters
        own = rng.integers(0, 400, n)
        a.sort(key)
        b.sort(key, own)
        # between the sorts the slot-bound state changes per PARCEL-IN-SLOT (the mesoscale wind memory)
        bump = rng.normal(size=n)                   # indexed by slot
        a.state = a.state + bump
        b.state = b.state + (bump if b.slot is None else bump[b.slot])
        ia, sa = a.by_slot()
        ib, sb = b.by_slot()
        assert np.array_equal(ia, ib) and np.array_equal(sa, sb), event
        assert np.all(np.diff(own[b.ident]) >= 0)   # the parcels really lie in the other key's order
        if event % 4 == 3:                          # a read-back in between restores position = slot for good
            b.unscramble()
            assert np.array_equal(b.ident, a.ident) and np.array_equal(b.state, a.state)
