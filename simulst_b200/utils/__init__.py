"""Mirror of the reference's ``codebase/utils`` package (hot-path functions only)."""
