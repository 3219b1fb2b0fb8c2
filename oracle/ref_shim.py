"""
ORACLE SIDE (test / bench infrastructure, not product code): import shim for the UNMODIFIED
reference (mitrefireline/simfire).

Two users: ``tests/golden/gen_*.py`` (golden vectors, dev container, reads ``/root/reference``)
and ``oracle/reference_runner.py`` (the CPU arm of ``bench.py`` and the same-run parity checks),
which imports the staged copy ``oracle/_ref/`` that ``oracle/build_ref.py`` makes from
``/root/reference`` -- git-ignored, so no reference source enters the history, but not
gpurun-ignored, so it travels to the GPU box.  Nothing in the product (``simfire_b200/``) imports it.

The reference is pure Python but imports ~12 display / GIS modules that are absent here
(pygame, matplotlib, reportlab, ...).  None of them is touched by the headless hot path
(``simfire/game/managers/fire.py:616-719`` -> ``simfire/world/rothermel.py:4-136``), so we
register inert stand-ins in ``sys.modules`` before importing.  The only stand-in with
behaviour is ``pygame.Rect`` (the reference keeps a sprite's (x, y) in ``Fire.rect``,
``simfire/game/sprites.py:223-227``, and unpacks it at ``fire.py:139``).
"""
from __future__ import annotations

import collections
import collections.abc
import importlib.metadata
import sys
import types

import os

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_ROOT = os.path.join(HERE, "_ref")  # oracle/build_ref.py


def reference_root() -> str:
    """Where the unmodified reference package lives: $SFB_REFERENCE_ROOT, the mounted checkout in the
    dev container, else the staged copy (the only one that exists on the GPU box)."""
    for cand in (os.environ.get("SFB_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.exists(os.path.join(cand, "simfire", "game", "managers", "fire.py")):
            return cand
    raise ImportError("the reference is neither mounted at /root/reference nor staged under oracle/_ref "
                      "(run `python oracle/build_ref.py` in the dev container)")


def available() -> bool:
    try:
        reference_root()
        return True
    except ImportError:
        return False


class _Rect:
    """Minimal pygame.Rect: x, y, w, h, iterable, move()."""

    def __init__(self, x, y=None, w=0, h=0):
        if y is None:  # Rect((x, y, w, h))
            x, y, w, h = x
        self.x, self.y, self.w, self.h = int(x), int(y), int(w), int(h)

    def __iter__(self):
        return iter((self.x, self.y, self.w, self.h))

    def move(self, dx, dy):
        return _Rect(self.x + dx, self.y + dy, self.w, self.h)

    def update(self, *a, **k):  # pragma: no cover
        pass


class _Sprite:
    def __init__(self, *a, **k):
        pass


class _Surface:
    """What `pygame.surfarray.make_surface(array)` returns, as far as `Fire.__init__` (sprites.py:229-236)
    uses it when a manager is not headless (ConstantSpreadFireManager cannot be, fire.py:752)."""

    def __init__(self, array):
        self.w, self.h = int(array.shape[0]), int(array.shape[1])

    def get_rect(self):
        return _Rect(0, 0, self.w, self.h)


class _Anything(types.ModuleType):
    """Module whose every attribute is an inert callable/class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _inert(*a, **k):
            return None

        return _inert


def _mod(name: str, **attrs) -> types.ModuleType:
    m = _Anything(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False


def install() -> None:
    """Register the stand-ins and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    _installed = True

    # Python >= 3.10 removed collections.Sequence (used at fire.py:415)
    if not hasattr(collections, "Sequence"):
        collections.Sequence = collections.abc.Sequence  # type: ignore[attr-defined]

    # simfire/__init__.py:27 asks importlib.metadata for its own version
    _orig_version = importlib.metadata.version

    def _version(name):
        if name == "simfire":
            return "2.0.1"
        return _orig_version(name)

    importlib.metadata.version = _version  # type: ignore[assignment]

    pg = _mod("pygame", Rect=_Rect)
    pg.rect = _mod("pygame.rect", Rect=_Rect)
    pg.sprite = _mod("pygame.sprite", Sprite=_Sprite)
    pg.surface = _mod("pygame.surface", Surface=object)
    pg.surfarray = _mod("pygame.surfarray", make_surface=_Surface)
    pg.display = _mod("pygame.display")
    pg.image = _mod("pygame.image")
    pg.transform = _mod("pygame.transform")
    pg.time = _mod("pygame.time")
    pg.event = _mod("pygame.event")
    pg.draw = _mod("pygame.draw")
    pg.font = _mod("pygame.font")

    mpl = _mod("matplotlib")
    mpl.pyplot = _mod("matplotlib.pyplot", Figure=object)
    mpl.lines = _mod("matplotlib.lines")
    mpl.contour = _mod("matplotlib.contour", QuadContourSet=object)

    rl = _mod("reportlab")
    rl.graphics = _mod("reportlab.graphics", renderPM=None)
    sv = _mod("svglib")
    sv.svglib = _mod("svglib.svglib")
    _mod("wurlitzer")
    _mod("geopandas")
    lf = _mod("landfire")
    lf.product = _mod("landfire.product")
    lf.product.enums = _mod(
        "landfire.product.enums", ProductRegion=object, ProductTheme=object, ProductVersion=object
    )
    lf.product.search = _mod("landfire.product.search", ProductSearch=object)
    gp = _mod("geopy")
    gp.distance = _mod("geopy.distance")
    _mod("geotiff", GeoTiff=object)
    _mod("h5py")
    _mod("jsonlines")
    _mod("noise")

    root = reference_root()
    if root not in sys.path:
        sys.path.insert(0, root)


def import_reference():
    """Return (fire_module, rothermel_module, enums, parameters, presets)."""
    install()
    from simfire import enums  # noqa: E402
    from simfire.game.managers import fire  # noqa: E402
    from simfire.world import parameters, presets, rothermel  # noqa: E402

    return fire, rothermel, enums, parameters, presets
