class PathInfo:
    pass


_VALID_CONTRACT_KWARGS = set()
