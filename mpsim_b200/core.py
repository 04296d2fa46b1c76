class MPS: pass
class MPSOperation: pass
class CannotConvertToMPSOperation(Exception): pass
