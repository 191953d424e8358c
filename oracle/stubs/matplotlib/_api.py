def check_in_list(values, /, *, _print_supported_values=True, **kwargs):
    for key, val in kwargs.items():
        if val not in values:
            raise ValueError(f"{val!r} is not a valid value for {key}; supported values are {values}")


def check_isinstance(types, /, **kwargs):
    pass
