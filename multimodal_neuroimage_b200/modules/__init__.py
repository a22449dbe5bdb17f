"""Host-side mirror of the reference's `modules/` package (same file and class names)."""
