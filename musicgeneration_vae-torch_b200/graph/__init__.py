"""B200-native mirror of the reference's ``graph`` package: same module tree and state_dict keys, compute in
libbarvae.so."""
