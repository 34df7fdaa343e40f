"""Mirror of the hot-path pieces of ``codebase/models`` of the reference."""
