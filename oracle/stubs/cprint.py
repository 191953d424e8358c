def c_print(*args, **kwargs):
    pass
