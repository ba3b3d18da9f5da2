# What the next GPU session should run first (written after round 2's GPU minutes were spent): the suite, then the measurements and
# checks of the code that was validated on the CPU only -- lock-step minimiser (Python and C++), random torsion trees on the device
T=${1:-r6a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^$" | cut -c1-400 | tail -40 > gpurun_out/${T}_pytest.log
python tests/device_tree_check.py > gpurun_out/${T}_tree_check.json 2> gpurun_out/${T}_tree_check.err
python tools/minimize_demo.py 1000 200 > gpurun_out/${T}_minimize.json 2> gpurun_out/${T}_minimize.err
./tests/cpp/host_test --minimize gnina_b200/weights > gpurun_out/${T}_cpp_minimize.log 2>&1
./tests/cpp/host_test --dock gnina_b200/weights > gpurun_out/${T}_cpp_dock.log 2>&1
./oracle/_ref/e2e_docking gpu > gpurun_out/${T}_e2e_docking.log 2>&1   # the reference's classes + the adapters + the device in one program
tail -3 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_tree_check.json gpurun_out/${T}_minimize.json gpurun_out/${T}_cpp_minimize.log gpurun_out/${T}_cpp_dock.log gpurun_out/${T}_e2e_docking.log
