# e2e piece schedules for rtb_render (f32 frame home): RTB_PIECE_WEIGHTS sets the relative piece sizes
for cfg in "X=1" "RTB_PIECE_WEIGHTS=5,8,8,7,6,5,3,2" "RTB_PIECE_WEIGHTS=4,7,8,7,6,5,4,3,2" "RTB_PIECE_WEIGHTS=6,9,8,7,6,5,4,3,2,1" "RTB_PIECE_WEIGHTS=3,6,8,8,7,6,5,4,3,2" "RTB_PIECE_WEIGHTS=9,8,7,6,5,4,3,2 RTB_LANES=3" "RTB_PIECES=6" "RTB_PIECES=10" "RTB_PIECES=12 RTB_LANES=3"; do
  echo "== $cfg"; env $cfg python tools/e2e_probe.py 2>&1 | grep rtb_render
done
