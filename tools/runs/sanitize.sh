for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E "smoke ok|mlp ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Invalid|Uninit" | head -20
done
