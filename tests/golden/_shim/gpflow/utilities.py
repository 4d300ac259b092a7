def positive():
    return None


def set_trainable(*a, **k):
    return None
