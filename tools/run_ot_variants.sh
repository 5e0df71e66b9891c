LIB=orb-slam-birdview_b200/liborbb200.so
cp $LIB /tmp/orig.so
for v in variants/*.so; do cp $v $LIB; echo "$(basename $v) $(tools/ubench/call_timeline 752 480 1000 300) $(tools/ubench/call_timeline 1241 376 2000 300)"; done
cp /tmp/orig.so $LIB
