"""Drop-in mirror of the reference's ``net`` package for the TEST path (B200 CUDA engine underneath)."""
