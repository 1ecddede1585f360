"""B200-native batched collision pipeline: GJK3/EPA3 narrow phase + SweepAndPrune3
broad phase of rustgd/collision-rs, behind the C ABI in include/collision_b200.h."""
