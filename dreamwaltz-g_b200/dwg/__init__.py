"""dwg: host-side mirror of the DreamWaltz-G SDS hot path on B200."""
