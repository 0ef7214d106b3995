# Warp-specialised K12 against the single-warp form: (runs, lanes, ref_frame) sweeps, both settings.
for spec in 1 0; do
  for cfg in "1000 16 1" "1000 8 1" "1000 32 1" "2000 16 1" "2000 8 1" "4000 8 1" "1000 16 0" "500 32 1"; do
    echo -n "{\"spec\": $spec, \"r\": "; B2INS_MC_SPEC=$spec python tools/probe_mc.py $cfg 20 | tr -d '\n'; echo "}"
  done
done
