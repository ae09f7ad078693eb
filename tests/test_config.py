"""ConfigParser semantics (config_parser.cpp:15-185, config_parser.hpp:131-151) through the C ABI."""
import json

import pytest

from opensayal_b200 import ConfigParser, SayalError
from opensayal_b200._abi import SAYAL_EIO, SAYAL_EPARSE


def parse(obj):
    return ConfigParser.parse_text(obj if isinstance(obj, str) else json.dumps(obj))


def test_defaults_match_reference():
    c = parse({}).c
    assert (c.width, c.height) == (1920, 1080)
    assert c.cell_size == 1.0 and c.enable_drain == 1 and c.enable_pressure == 0 and c.enable_smoke == 1
    assert c.proj_n == 50 and abs(c.proj_o - 1.9) < 1e-7
    assert c.wt_pipe_height == 270 and c.wt_smoke_height == 270 and c.wt_pipe_length == 0
    assert c.wt_smoke_length == 1 and c.wt_smoke_count == 1 and c.wt_speed == 0.0 and c.wt_smoke == 1.0
    assert c.g == 0.0 and abs(c.d_t - 0.05) < 1e-9
    assert c.smoke_enable_decay == 0 and abs(c.smoke_decay_rate - 0.05) < 1e-9
    assert c.obstacle_enable == 1 and (c.obstacle_center_x, c.obstacle_center_y) == (960, 540)
    assert abs(c.obstacle_radius - 36.0) < 1e-6
    assert c.density == 1.0 and c.drag_coeff == 0.0 and abs(c.viscosity - 0.001) < 1e-9
    assert (c.block_size_x, c.block_size_y) == (64, 1)


def test_defaults_follow_dimensions():
    c = parse({"sim": {"width": 300, "height": 200}}).c
    assert c.wt_pipe_height == 50 and c.wt_smoke_height == 50
    assert (c.obstacle_center_x, c.obstacle_center_y) == (150, 100)
    assert abs(c.obstacle_radius - 200 / 30.0) < 1e-6


@pytest.mark.parametrize("doc", [
    {"sim": {"time": {"d_t": 0.125}}},
    {"sim.time.d_t": 0.125},
    {"sim": {"time.d_t": 0.125}},
])
def test_three_equivalent_spellings(doc):  # README.md:162-186
    assert parse(doc).c.d_t == 0.125


def test_split_only_at_first_dot():  # H15: {"sim.time": {...}} is NOT recognised
    assert abs(parse({"sim.time": {"d_t": 0.125}}).c.d_t - 0.05) < 1e-9


def test_literal_key_wins_over_nested():
    assert parse({"sim.time.d_t": 0.25, "sim": {"time": {"d_t": 0.5}}}).c.d_t == 0.25


def test_int_for_float_and_float_for_int():
    c = parse({"sim": {"cell_size": 1, "wind_tunnel": {"speed": 200}, "projection": {"n": 50.9}}}).c
    assert c.cell_size == 1.0 and c.wt_speed == 200.0 and c.proj_n == 50


def test_nested_type_error_is_swallowed_top_level_is_fatal():
    # nested: the parent's catch(...) returns the default (config_parser.hpp:145-150)
    assert parse({"sim": {"enable_drain": 7}}).c.enable_drain == 1
    assert parse({"sim": {"projection": {"n": "many"}}}).c.proj_n == 50
    # top-level literal key: nlohmann's type_error propagates in the reference; here SAYAL_EPARSE
    with pytest.raises(SayalError) as e:
        parse({"sim.enable_drain": 7})
    assert e.value.code == SAYAL_EPARSE


def test_parent_not_an_object_gives_default():
    assert parse({"sim": 3}).c.width == 1920
    assert parse({"sim": {"time": 4}}).c.d_t == pytest.approx(0.05)


def test_sample_config_shape():
    doc = {"thread": {"cuda": {"block_size_x": 64, "block_size_y": 1}},
           "sim": {"height": 1080, "width": 1920, "cell_size": 1, "enable_drain": True, "enable_pressure": False,
                   "enable_smoke": False, "projection": {"n": 50, "o": 1.9}, "enable_interactive": True,
                   "wind_tunnel": {"pipe_height": 1080, "smoke_length": 1, "speed": 200, "smoke": 1},
                   "physics": {"g": 0}, "time": {"d_t": 0.05, "enable_read_time": False},
                   "smoke": {"enable_decay": False, "decay_rate": 0},
                   "obstacle": {"enable": True, "center_x": 960, "center_y": 540, "radius": 36}},
           "fluid": {"density": 1}, "visual": {"arrows": {"enable": False}}}
    c = parse(doc).c
    assert c.enable_smoke == 0 and c.enable_interactive == 1 and c.wt_pipe_height == 1080
    assert c.wt_speed == 200.0 and c.smoke_decay_rate == 0.0 and c.obstacle_radius == 36.0
    assert c.enable_real_time == 0  # "enable_read_time" is a typo in the shipped sample and is ignored


def test_errors(tmp_path):
    with pytest.raises(SayalError) as e:
        ConfigParser(str(tmp_path / "missing.json")).parse()
    assert e.value.code == SAYAL_EIO
    for bad in ("{", '{"a": }', '{"a": 1,}', "[1, 2", '{"a": tru}', '{"a": 1} x'):
        with pytest.raises(SayalError) as e:
            parse(bad)
        assert e.value.code == SAYAL_EPARSE
    p = tmp_path / "OpenSayal.conf.json"
    p.write_text(json.dumps({"sim": {"width": 640, "height": 360, "physics.g": -9.5}}))
    c = ConfigParser(str(p)).parse().c
    assert (c.width, c.height, c.g) == (640, 360, -9.5)


def test_config_views():
    cfg = parse({"sim": {"projection": {"n": 7}}})
    assert cfg.sim.projection.n == 7 and cfg["sim.projection.n"] == 7
    cfg.sim.projection.n = 9
    assert cfg.c.proj_n == 9
    assert cfg.fluid.density == 1.0


def test_python_restatement_of_the_defaults_matches_the_library():
    """bench.py's reference arm builds its configuration from oracle.reference_defaults (so that it never maps the
    product library); it must be the same struct, byte for byte, as sayal_config_defaults gives."""
    import ctypes as C

    from opensayal_b200 import load
    from opensayal_b200._abi import SayalConfig
    from opensayal_b200.synthetic import baseline_config
    from oracle.oracle import reference_defaults
    for w, h in ((1920, 1080), (256, 144), (101, 67), (16384, 16384), (3840, 2160)):
        c = SayalConfig()
        assert load().sayal_config_defaults(w, h, C.byref(c)) == 0
        assert bytes(c) == bytes(reference_defaults(w, h)), (w, h)
    for index in range(5):
        assert bytes(baseline_config(index).c) == bytes(baseline_config(index, defaults=reference_defaults).c)
