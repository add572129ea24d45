class DeepDiff(dict):
    def __init__(self, *a, **k):
        super().__init__()
