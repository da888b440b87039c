"""Drop-in package: the reference's `nff.*` names that the MD hot path imports, backed by mdgrad_b200."""
