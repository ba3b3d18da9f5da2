# Short GPU check of one change: the GPU test-suite, the smoke call and the docking-kernel rates (outputs under gpurun_out/)
T=${1:-r5a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-400 | tail -60 > gpurun_out/${T}_pytest.log
python tools/dock_rows.py 4096 40 > gpurun_out/${T}_dock_rows.json 2> gpurun_out/${T}_dock_rows.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1
tail -5 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_dock_rows.json; tail -2 gpurun_out/${T}_smoke.log
