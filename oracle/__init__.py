"""TEST INFRASTRUCTURE — CPU oracle for the vectorized env step.

This package is a numpy restatement of the reference's algorithm for the hot
path: ``TaxiVecEnv.step``, ``RoomsEnv.step``, the observation
functions, ``CRoomsEnv.step`` and the ant-tag pursuit rules.  Every function
cites the reference file:line it follows.

It is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package (``gym-po-taxi_b200/``)
never does and fails loudly when its CUDA library is missing.

Parity pinning: PINNED.  ``tests/golden/*.npz`` hold trajectories produced by
executing the real reference (``/root/reference`` loaded through
``oracle.ref_loader``, script ``tests/golden/make_golden.py``); the ``not gpu``
tests replay the recorded random draws through this oracle and require
bit-identical obs / reward / terminated / truncated / state, and — when the
reference directory is present — also step the oracle and the reference side by
side from the same seed.

Random numbers: each oracle env pulls its randomness from a *draw source*
(``oracle.draws``).  ``GeneratorDraws`` issues exactly the reference's
``np.random.Generator`` calls in the reference's order (so equal seeds give equal
trajectories); ``RecordedDraws`` replays values captured from the reference.  After
every ``reset``/``step`` the env exposes ``draws``: dense per-env arrays of the
values consumed, in the layout the CUDA library's replay mode reads
(SURVEY.md Appendix B).
"""
from .draws import GeneratorDraws, RecordedDraws, make_generator  # noqa: F401
from .taxi import TaxiOracle, TAXI_MAP, EXTENDED_TAXI_MAP  # noqa: F401
from .rooms import RoomsOracle, load_layout, LAYOUT_NAMES  # noqa: F401
from .crooms import CRoomsOracle  # noqa: F401
from .tag import tag_move_target, TagOracle  # noqa: F401
from .car import CarOracle  # noqa: F401
from .msrooms import MSRoomsOracle  # noqa: F401
from .wrappers import RecordEpisodeStatisticsOracle, NormalizeRewardOracle, RunningMeanStd  # noqa: F401
