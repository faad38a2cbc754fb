"""Host mirrors of the reference's ``processing`` package for the scoring pass."""
