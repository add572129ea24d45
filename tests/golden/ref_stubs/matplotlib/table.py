class Table:  # placeholder, never instantiated on the hot path
    pass
