"""Pick the checker for a model: oracle/_ref (the unmodified reference C, prebuilt .so or built from
/root/reference) when available, else the C restatement in oracle/ (kind 'port')."""
from oracle import ref


class _Ref:
    kind = "reference"

    def __init__(self, model):
        self.r = ref.Reference(model)

    def solve(self):
        return self.r.solve()

    def simulate(self, M, D, init, rs, rndtype=0):
        return self.r.simulate(M, D, init, rs, rndtype)

    @property
    def seconds(self):
        return self.r.last_seconds


class _Port:
    kind = "port"

    def __init__(self, model):
        from oracle import port
        self.p = port.Port(model)

    def solve(self):
        return self.p.solve()  # raises: the solver's checker is the compiled reference only

    def simulate(self, M, D, init, rs, rndtype=0):
        return self.p.simulate(M, D, init, rs, rndtype)

    @property
    def seconds(self):
        return self.p.last_seconds


def ref_available(model) -> bool:
    return ref.build(model) is not None


def oracle_for(model, prefer="reference"):
    model.prepare()
    if prefer == "reference" and ref_available(model):
        return _Ref(model)
    return _Port(model)
