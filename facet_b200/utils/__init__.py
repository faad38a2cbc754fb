"""Host-side helpers mirroring the reference's ``utils`` package for the scoring pass."""
