class JVPTracer:  # never instantiated by the stand-in
    pass
