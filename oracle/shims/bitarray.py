"""Minimal stand-in for the `bitarray` package (not installed in this image), covering exactly the
subset the reference uses (entropy_encoder.py:18-27,35-60, PFrame.py:142-175, Frame.py:62-87):
bitarray(), bitarray(n) -> n zero bits (bitarray>=3.0 semantics, requirements.txt:7), bitarray(str),
frombytes/tobytes (big-endian bit order, zero padded), extend, index, slicing, +, len, truthiness,
to01, integer indexing.  TEST INFRASTRUCTURE ONLY -- used to import the Python reference."""


class bitarray:
    __slots__ = ("_b",)

    def __init__(self, init=None):
        if init is None:
            self._b = []
        elif isinstance(init, int):
            self._b = [0] * init
        elif isinstance(init, str):
            self._b = [1 if c == "1" else 0 for c in init]
        elif isinstance(init, bitarray):
            self._b = list(init._b)
        else:
            self._b = [1 if x else 0 for x in init]

    def frombytes(self, data):
        ext = self._b.extend
        for byte in bytes(data):
            ext(((byte >> 7) & 1, (byte >> 6) & 1, (byte >> 5) & 1, (byte >> 4) & 1,
                 (byte >> 3) & 1, (byte >> 2) & 1, (byte >> 1) & 1, byte & 1))

    def tobytes(self):
        b = self._b
        n = len(b)
        out = bytearray((n + 7) // 8)
        for i in range(n):
            if b[i]:
                out[i >> 3] |= 0x80 >> (i & 7)
        return bytes(out)

    def extend(self, other):
        if isinstance(other, bitarray):
            self._b.extend(other._b)
        else:
            self._b.extend(1 if x else 0 for x in other)

    def append(self, x):
        self._b.append(1 if x else 0)

    def index(self, value, *a):
        return self._b.index(1 if value else 0, *a)

    def to01(self):
        return "".join("1" if x else "0" for x in self._b)

    def __len__(self):
        return len(self._b)

    def __bool__(self):
        return len(self._b) > 0

    def __getitem__(self, k):
        if isinstance(k, slice):
            r = bitarray()
            r._b = self._b[k]
            return r
        return self._b[k]

    def __add__(self, other):
        r = bitarray()
        r._b = self._b + other._b
        return r

    def __eq__(self, other):
        return isinstance(other, bitarray) and self._b == other._b

    def __iter__(self):
        return iter(self._b)

    def __repr__(self):
        return f"bitarray('{self.to01()}')"
