#!/usr/bin/env bash
# Installs the UNMODIFIED reference (qibo 0.3.5, pure Python) into baseline/_ref (git-ignored).
# 1) tries the contract's offline pip install; 2) if the build backend (poetry-core) is missing from the
# wheelhouse, performs the equivalent of a pure-Python wheel install: package dir + dist-info metadata.
# Also drops a stub `openqasm3` module (reference imports it eagerly; it is absent from this image).
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
DST="$HERE/_ref"
[ -d "$REF/src/qibo" ] || { echo "no reference at $REF"; exit 0; }
mkdir -p "$DST"
if python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
      --target "$DST" "$REF" >/tmp/qibo_pip.log 2>&1; then
  echo "pip install ok"
else
  echo "pip install failed (see /tmp/qibo_pip.log); doing a plain pure-python install"
  rm -rf "$DST/qibo"
  cp -r "$REF/src/qibo" "$DST/qibo"
  mkdir -p "$DST/qibo-0.3.5.dist-info"
  printf 'Metadata-Version: 2.1\nName: qibo\nVersion: 0.3.5\n' > "$DST/qibo-0.3.5.dist-info/METADATA"
  printf 'manual\n' > "$DST/qibo-0.3.5.dist-info/INSTALLER"
fi
# The reference's OWN test-suite, for running it against the qibo_b200 backend on the GPU box (scripts/run_reference_tests.sh).
# Same git-ignored directory; the only edit is to its conftest.py: our backend is added to the BACKENDS list and a missing
# qibojit (MissingBackend is a ValueError, backends/__init__.py:17,346) is tolerated.
if [ -d "$REF/tests" ]; then
  rm -rf "$DST/tests"
  mkdir -p "$DST/tests"
  cp "$REF"/tests/__init__.py "$REF"/tests/conftest.py "$DST/tests/"
  [ -f "$REF/tests/utils.py" ] && cp "$REF/tests/utils.py" "$DST/tests/"
  for t in test_gates_gates test_gates_density_matrix test_gates_special test_gates_abstract test_backends test_backends_global \
           test_models_circuit_fuse test_models_qft test_models_circuit_execution test_models_circuit_features \
           test_models_circuit_parametrized test_models_circuit test_measurements test_measurements_probabilistic \
           test_measurements_collapse test_result test_states test_models_distcircuit test_models_distcircuit_execution \
           test_callbacks test_parallel test_models_circuit_noise test_hamiltonians_terms test_models_variational \
           test_models_evolution test_hamiltonians_trotter test_hamiltonians_symbolic test_hamiltonians_models \
           test_gates_channels test_noise test_derivative test_models_grover test_models_encodings; do
    [ -f "$REF/tests/$t.py" ] && cp "$REF/tests/$t.py" "$DST/tests/"
  done
  [ -d "$REF/tests/regressions" ] && cp -r "$REF/tests/regressions" "$DST/tests/"
  sed -i 's/^    "qibojit-numba",$/    "qibo_b200",/; s/^    except ImportError:$/    except (ImportError, ValueError):/; s/^except ImportError:$/except (ImportError, ValueError):/' "$DST/tests/conftest.py"
fi
python -c "import openqasm3" 2>/dev/null || { mkdir -p "$DST/openqasm3"; printf 'parser = None\nast = None\n' > "$DST/openqasm3/__init__.py"; }
PYTHONPATH="$DST" python -c "import qibo; print('qibo', qibo.__version__, 'importable from', qibo.__file__)"
