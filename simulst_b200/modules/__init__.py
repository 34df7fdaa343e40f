"""Mirror of the hot-path method bodies of ``codebase/modules`` of the reference."""
