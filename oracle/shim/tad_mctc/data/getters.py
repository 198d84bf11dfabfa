def get_zvalence(numbers, device=None, dtype=None):
    raise NotImplementedError("get_zvalence is not carried by the oracle shim (dipole moments are outside the hot path)")


def get_atomic_masses(numbers, atomic_units=True, device=None, dtype=None):
    raise NotImplementedError("atomic masses are not carried by the oracle shim")
