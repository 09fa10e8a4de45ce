"""Mirror of the reference's `utils` package for the box-level hot path."""
