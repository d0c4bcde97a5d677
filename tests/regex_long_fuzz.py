"""Long differential run of the regexp engine (csrc/host_regex.hpp through tfgpu_regex_replace_all) against oracle/regex_oracle.py:
30 seeds x 4000 random expressions, about 350 k expression / text pairs (the test suite runs one seed). python tests/regex_long_fuzz.py (not collected by pytest)"""
import os
import random
import signal
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_regex_replace as T
from oracle import regex_oracle as ro
from transferia_b200 import engine, sink
R=sink.regex_replace_all
alphabet = ["a", "b", "c", "ab", "1", "_", " ", "\n", ".", "-", "日", "é", b"\xff", b"\xe6\x97", "x", "A", "B", "k", "K", "S", "s", "K", "ſ"]
rules = ["", "X", "<$0>", "[$1|$2]", "${1}x$1x", "$$1", "$", "${g}", "$g1234", "a$0b$9"]
class OracleTooSlow(Exception):
    pass


def _alarm(_sig, _frm):
    raise OracleTooSlow()                  # Python's backtracking engine goes exponential on some nested repeats; the product's machine is linear


signal.signal(signal.SIGALRM, _alarm)
bad=0; n=0; skipped=0
for seed in range(1, 31):
    rng = random.Random(seed*7919)
    for _ in range(4000):
        pat,_n = T._gen(rng); rule=rng.choice(rules)
        try: ro.compile_go(pat)
        except (ro.GoSyntaxError, NotImplementedError):
            try: R(pat, rule, b""); print("PRODUCT ACCEPTS", repr(pat)); bad+=1
            except engine.EngineError: pass
            continue
        for _ in range(3):
            src = b"".join(x if isinstance(x, bytes) else x.encode() for x in (rng.choice(alphabet) for _ in range(rng.randrange(0, 14))))
            if not src and "\\B" in pat: continue
            try:
                got = R(pat, rule, src)
            except engine.EngineError as e:
                print("PRODUCT REFUSES", repr(pat), e.rc); bad+=1; break
            signal.alarm(5)
            try: want = ro.replace_all(pat, rule, src); n+=1
            except OracleTooSlow: skipped+=1; break
            finally: signal.alarm(0)
            if got != want: print("DIFF", repr(pat), repr(rule), src, got, want); bad+=1
    print("seed", seed, "checked", n, "bad", bad, "oracle too slow", skipped, flush=True)
    if bad > 20: break
