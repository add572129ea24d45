class GMM:  # imported by the reference, never used on the MaxScoreBatchSubsetWithSkips path
    pass
