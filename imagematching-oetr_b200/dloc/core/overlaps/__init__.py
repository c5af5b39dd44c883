def get_overlap(overlaper):
    """reference dloc/core/overlaps/__init__.py:10-12"""
    mod = __import__(f'{__name__}.{overlaper}', fromlist=[''])
    return getattr(mod, 'Model')
