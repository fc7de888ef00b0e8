//! parity_dump.rs -- writes the golden vectors that pin the B200 chunk pipeline to THIS crate.
//!
//! The CUDA path (and its CPU oracle) restate libnoise 1.1.2 from memory: no machine that built
//! them had the crate's source.  This test runs inside the reference and dumps what is needed to
//! close that gap: raw `Simplex` / `Fbm` samples, probes next to lattice corners (they reveal the
//! order of the crate's private gradient tables) and whole generated areas.
//!
//! How to run (any machine with cargo; nothing else of the game is touched):
//!   1. copy this file to  src/service/area_generation/parity_dump.rs
//!   2. add the line       #[cfg(test)] mod parity_dump;        to src/service/area_generation/mod.rs
//!   3. cargo test --release parity_dump -- --nocapture
//!   4. copy  parity_dump.json  (or $VOXEL_PARITY_DUMP) to  tests/golden/reference/dump.json  of the
//!      B200 repository and run  python -m pytest tests/test_reference_dump.py -s
//! The B200 test then reports which setting of the calibration switches (VX_NOISE_VARIANT,
//! vx_set_gradient_order) reproduces the dump bit for bit, or what differs first.
//!
//! f64 values are written as the decimal u64 of `to_bits()`: exact, and plain JSON.
use std::fmt::Write as _;
use std::hash::{DefaultHasher, Hash, Hasher};

use libnoise::{Generator, Simplex};

use crate::model::{
    area::{AREA_HEIGHT, AREA_SIZE},
    location::{AreaLocation, InternalLocation},
};
use crate::service::area_generation::generator::AreaGenerator;

const WORLD_NAME: &str = "test";
const UNSKEW_2D: f64 = 0.211324865405187;
const UNSKEW_3D: f64 = 1.0 / 6.0;
const PROBE_H: f64 = 0.01;

fn hash_world_name(world_name: &str) -> u64 {
    // generator.rs:24-28, verbatim (the function is private there)
    let mut hasher = DefaultHasher::new();
    world_name.hash(&mut hasher);
    hasher.finish()
}

/// SplitMix64: the points are arbitrary, they only have to be written down
struct Rng(u64);
impl Rng {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9e3779b97f4a7c15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xbf58476d1ce4e5b9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94d049bb133111eb);
        z ^ (z >> 31)
    }
    fn uniform(&mut self, lo: f64, hi: f64) -> f64 {
        lo + (hi - lo) * ((self.next() >> 11) as f64 / (1u64 << 53) as f64)
    }
}

fn bits_list(values: &[f64]) -> String {
    let mut s = String::from("[");
    for (i, v) in values.iter().enumerate() {
        if i > 0 {
            s.push(',');
        }
        write!(s, "{}", v.to_bits()).unwrap();
    }
    s.push(']');
    s
}

fn u32_list(values: &[u32]) -> String {
    let mut s = String::from("[");
    for (i, v) in values.iter().enumerate() {
        if i > 0 {
            s.push(',');
        }
        write!(s, "{}", v).unwrap();
    }
    s.push(']');
    s
}

fn hex(bytes: &[u8]) -> String {
    let mut s = String::with_capacity(bytes.len() * 2);
    for b in bytes {
        write!(s, "{:02x}", b).unwrap();
    }
    s
}

fn sample_points(rng: &mut Rng, dim: usize) -> Vec<f64> {
    let mut pts = Vec::new();
    // small coordinates, the scale of the cave noise at world coordinates (1e6 * 0.08), the raw
    // world coordinates, negative values, exact ties (x == y [== z]) and exact lattice points
    for (lo, hi, n) in [(-4.0, 4.0, 300), (79_000.0, 81_000.0, 300), (999_000.0, 1_001_000.0, 200), (-2_000.0, 0.0, 100)] {
        for _ in 0..n {
            for _ in 0..dim {
                pts.push(rng.uniform(lo, hi));
            }
        }
    }
    for _ in 0..100 {
        let v = rng.uniform(-50.0, 50.0);
        for _ in 0..dim {
            pts.push(v);
        }
    }
    for i in 0..50usize {
        for d in 0..dim {
            pts.push((i * (d + 1)) as f64);
        }
    }
    pts
}

#[test]
fn parity_dump() {
    let seed = hash_world_name(WORLD_NAME);
    let seeds = [seed, seed ^ u64::MAX, seed.wrapping_add(1), seed.wrapping_add(10)];
    let mut rng = Rng(0x5EED_0001);
    let mut out = String::new();
    write!(out, "{{\"format\":1,\"crate\":\"voxel_game 1.7.0 / libnoise 1.1.2\",\"world_name\":\"{}\",\"seed\":{},", WORLD_NAME, seed).unwrap();

    // ---- raw simplex samples + corner probes
    out.push_str("\"simplex\":[");
    let mut first = true;
    for &s in &seeds {
        // 2-D
        let g2: Simplex<2> = Simplex::new(s);
        let pts = sample_points(&mut rng, 2);
        let vals: Vec<f64> = pts.chunks(2).map(|p| g2.sample([p[0], p[1]])).collect();
        if !first {
            out.push(',');
        }
        first = false;
        write!(out, "{{\"seed\":{},\"dim\":2,\"points\":{},\"values\":{}}}", s, bits_list(&pts), bits_list(&vals)).unwrap();
        // probes: lattice corner (i, j) unskewed, then h along each axis.  Next to a corner only that
        // corner is inside its radius, so value = NORM * (0.5 - h^2)^4 * h * gradient[axis].
        let (mut lattice, mut ppts) = (Vec::new(), Vec::new());
        for i in 0..16u32 {
            for j in 0..16u32 {
                let un = (i + j) as f64 * UNSKEW_2D;
                let (cx, cy) = (i as f64 - un, j as f64 - un);
                for axis in 0..2 {
                    lattice.extend_from_slice(&[i, j, axis]);
                    ppts.push(cx + if axis == 0 { PROBE_H } else { 0.0 });
                    ppts.push(cy + if axis == 1 { PROBE_H } else { 0.0 });
                }
            }
        }
        let pvals: Vec<f64> = ppts.chunks(2).map(|p| g2.sample([p[0], p[1]])).collect();
        write!(out, ",{{\"seed\":{},\"dim\":2,\"probe_h\":{},\"lattice\":{},\"points\":{},\"values\":{}}}", s,
               PROBE_H.to_bits(), u32_list(&lattice), bits_list(&ppts), bits_list(&pvals)).unwrap();
        // 3-D
        let g3: Simplex<3> = Simplex::new(s);
        let pts = sample_points(&mut rng, 3);
        let vals: Vec<f64> = pts.chunks(3).map(|p| g3.sample([p[0], p[1], p[2]])).collect();
        write!(out, ",{{\"seed\":{},\"dim\":3,\"points\":{},\"values\":{}}}", s, bits_list(&pts), bits_list(&vals)).unwrap();
        let (mut lattice, mut ppts) = (Vec::new(), Vec::new());
        for i in 0..7u32 {
            for j in 0..7u32 {
                for k in 0..7u32 {
                    let un = (i + j + k) as f64 * UNSKEW_3D;
                    let c = [i as f64 - un, j as f64 - un, k as f64 - un];
                    for axis in 0..3 {
                        lattice.extend_from_slice(&[i, j, k, axis]);
                        for d in 0..3 {
                            ppts.push(c[d] + if axis as usize == d { PROBE_H } else { 0.0 });
                        }
                    }
                }
            }
        }
        let pvals: Vec<f64> = ppts.chunks(3).map(|p| g3.sample([p[0], p[1], p[2]])).collect();
        write!(out, ",{{\"seed\":{},\"dim\":3,\"probe_h\":{},\"lattice\":{},\"points\":{},\"values\":{}}}", s,
               PROBE_H.to_bits(), u32_list(&lattice), bits_list(&ppts), bits_list(&pvals)).unwrap();
    }
    out.push_str("],");

    // ---- the nine Fbm objects of the generator at world coordinates (terrain_type.rs:16-18,
    // biome_type.rs:21, cave_generator.rs:27-32, voxel_type_generator.rs:27, lake_generator.rs:22)
    let fbms: [(&str, u64, usize, u32, f64, f64, f64); 9] = [
        ("height1", seed, 2, 6, 0.008, 1.8, 0.45),
        ("height2", seed ^ u64::MAX, 2, 4, 0.004, 1.8, 0.45),
        ("height_mod", seed.wrapping_add(1), 2, 2, 0.004, 0.004, 0.2),
        ("biome", seed, 2, 2, 0.0015, 1.5, 0.3),
        ("cave1", seed, 3, 6, 0.08, 1.3, 0.45),
        ("cave2", seed ^ u64::MAX, 3, 8, 0.04, 1.2, 0.35),
        ("cave_check", seed, 2, 2, 0.003, 1.3, 0.3),
        ("alt", seed.wrapping_add(1), 3, 2, 0.07, 2.0, 0.3),
        ("lake", seed.wrapping_add(10), 2, 2, 0.004, 1.3, 0.35),
    ];
    out.push_str("\"fbm\":[");
    for (n, &(name, s, dim, octaves, f, l, p)) in fbms.iter().enumerate() {
        let mut pts = Vec::new();
        for _ in 0..400 {
            // get_point_on_noise_map: (x + area.x * 16) as f64 around the spawn area 62500, z_inverted as f64
            pts.push(((rng.next() % 40_000) + 980_000) as f64);
            pts.push(((rng.next() % 40_000) + 980_000) as f64);
            if dim == 3 {
                pts.push(((rng.next() % 127) + 1) as f64);
            }
        }
        let vals: Vec<f64> = if dim == 2 {
            let g = Simplex::<2>::new(s).fbm(octaves, f, l, p);
            pts.chunks(2).map(|q| g.sample([q[0], q[1]])).collect()
        } else {
            let g = Simplex::<3>::new(s).fbm(octaves, f, l, p);
            pts.chunks(3).map(|q| g.sample([q[0], q[1], q[2]])).collect()
        };
        if n > 0 {
            out.push(',');
        }
        write!(out, "{{\"name\":\"{}\",\"seed\":{},\"dim\":{},\"octaves\":{},\"frequency\":{},\"lacunarity\":{},\"persistence\":{},\"points\":{},\"values\":{}}}",
               name, s, dim, octaves, f.to_bits(), l.to_bits(), p.to_bits(), bits_list(&pts), bits_list(&vals)).unwrap();
    }
    out.push_str("],");

    // ---- whole areas: spawn, the reference's own test area (generator.rs:157), a 4x4 block with
    // caves and trees, far corners
    let mut areas: Vec<(u32, u32)> = vec![(62500, 62500), (123, 456), (999, 400), (0, 123), (70000, 1000)];
    for x in 62496..62500u32 {
        for y in 62502..62506u32 {
            areas.push((x, y));
        }
    }
    out.push_str("\"areas\":[");
    for (n, &(ax, ay)) in areas.iter().enumerate() {
        let area = AreaGenerator::generate_area(AreaLocation::new(ax, ay), WORLD_NAME);
        let mut voxels = vec![0u8; (AREA_SIZE * AREA_SIZE * AREA_HEIGHT) as usize];
        let mut heights = vec![0u8; (AREA_SIZE * AREA_SIZE) as usize];
        for z in 0..AREA_HEIGHT {
            for y in 0..AREA_SIZE {
                for x in 0..AREA_SIZE {
                    // the in-memory layout of Area (area.rs:58-64): x + 16 * y + 256 * z, z = 0 on top
                    voxels[(x + y * AREA_SIZE + z * AREA_SIZE * AREA_SIZE) as usize] =
                        area.get(InternalLocation::new(x, y, z)) as u8;
                }
            }
        }
        for y in 0..AREA_SIZE {
            for x in 0..AREA_SIZE {
                heights[(x + y * AREA_SIZE) as usize] = area.sample_height(x, y);
            }
        }
        if n > 0 {
            out.push(',');
        }
        write!(out, "{{\"x\":{},\"y\":{},\"voxels\":\"{}\",\"max_height\":\"{}\"}}", ax, ay, hex(&voxels), hex(&heights)).unwrap();
    }
    out.push_str("]}");

    let path = std::env::var("VOXEL_PARITY_DUMP").unwrap_or_else(|_| "parity_dump.json".to_string());
    std::fs::write(&path, out).expect("cannot write the dump");
    println!("parity dump written to {path}");
}
