"""Host-side mirror of the reference's `encoder` package for the hot path (same names and attributes)."""
