"""Mirror of the reference's `components` package for the hot path."""
