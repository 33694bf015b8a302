python tools/chamfer_step.py --kind uniform --kind2 blob --steps 1 --b 2 --no-backward 2>&1 | tail -3
compute-sanitizer --tool memcheck --print-limit 3 python tools/chamfer_step.py --kind uniform --kind2 blob --steps 1 --b 2 --no-backward 2>&1 | grep -v "Host Frame\|^=========$" | head -24
compute-sanitizer --tool racecheck --print-limit 3 python tools/chamfer_step.py --kind uniform --kind2 blob --steps 1 --b 1 --no-backward 2>&1 | grep -v "Host Frame\|^=========$" | head -24
