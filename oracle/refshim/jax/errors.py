class ConcretizationTypeError(TypeError):
    pass
