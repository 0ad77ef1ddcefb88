//! Known-answer vectors of crate B's numerics (see Cargo.toml).  One record per line, floats as IEEE-754 bit
//! patterns (hex).  The test image is a smooth function of a stateless 32-bit hash, values in [0, 1];
//! tests/test_crate_b_oracle.py reproduces it in numpy.
use feature_tracker::image_operations::{build_image_pyramid, d_interpolate_bicubic, interpolate_bicubic};
use feature_tracker::patch::{MatchingCost, Patch, Patch52};
use feature_tracker::types::{FloatGrayImage, RowVec2};
use nalgebra as na;
use std::fmt::Write as _;

fn mix(x: u32, y: u32, seed: u32) -> u32 {
    let mut s = x.wrapping_mul(0x9E37_79B1) ^ y.wrapping_mul(0x85EB_CA77) ^ seed.wrapping_mul(0xC2B2_AE3D);
    s ^= s >> 15;
    s = s.wrapping_mul(0x2C1B_3C6D);
    s ^= s >> 12;
    s = s.wrapping_mul(0x297A_2D39);
    s ^= s >> 15;
    s
}

/// 8-px blocks of hashed grey levels (k / 255, k in 0..=255) plus 4 bits of per-pixel noise: exactly representable inputs
fn scene(w: u32, h: u32, seed: u32) -> FloatGrayImage {
    FloatGrayImage::from_fn(w, h, |x, y| {
        let block = (mix(x / 8, y / 8, seed) >> 24) & 0xE0;
        let noise = mix(x, y, seed ^ 0x5555) >> 28;
        image::Luma([((block + noise) as f32) / 255.0])
    })
}

fn bits(x: f32) -> String {
    format!("{:08x}", x.to_bits())
}
fn dump(v: &[f32]) -> String {
    let mut s = String::with_capacity(9 * v.len());
    for x in v {
        write!(s, "{} ", bits(*x)).unwrap();
    }
    s.trim_end().to_string()
}

fn main() {
    let path = std::env::args().nth(1).unwrap_or_else(|| "reference_vectors_b.txt".to_string());
    let mut out = String::new();
    writeln!(out, "# rsvio reference vectors, crate B, v1").unwrap();
    for &(w, h, seed, levels, ratio, blur, sigma) in
        &[(120u32, 60u32, 61u32, 3usize, 2.0f64, false, 0.0f32), (97, 75, 62, 4, 2.0, false, 0.0), (96, 64, 63, 3, 1.5, false, 0.0), (120, 60, 61, 3, 2.0, true, 2.0)]
    {
        let img = scene(w, h, seed);
        let pyr = build_image_pyramid(&img, levels, ratio, blur, sigma);
        for (l, level) in pyr.iter().enumerate() {
            let (lw, lh) = level.dimensions();
            writeln!(out, "pyr_b {} {} {} {} {} {} {} {} {} {} : {}", w, h, seed, levels, bits(ratio as f32), blur as u8, bits(sigma), l, lw, lh, dump(level.as_raw())).unwrap();
        }
    }
    let img = scene(96, 72, 64);
    for &(x, y) in &[(10.0f32, 10.0f32), (20.25, 31.75), (47.5, 35.5), (80.125, 60.875), (0.5, 3.0), (93.2, 10.0)] {
        let mut g = RowVec2::zeros();
        let v = d_interpolate_bicubic(&[x, y], &img, &mut g);
        let v2 = interpolate_bicubic(&[x, y], &img);
        match (v, v2) {
            (Some(v), Some(v2)) => writeln!(out, "bicubic 96 72 64 {} {} 1 : {} {} {} {}", bits(x), bits(y), bits(v), bits(v2), bits(g[0]), bits(g[1])).unwrap(),
            _ => writeln!(out, "bicubic 96 72 64 {} {} 0 :", bits(x), bits(y)).unwrap(),
        }
    }
    for &(x, y) in &[(30.0f32, 30.0f32), (48.5, 36.25), (4.0, 4.0)] {
        let p = Patch52::new(&img, &na::Point2::new(x, y), 0.1, MatchingCost::SSD);
        let j: Vec<f32> = (0..52).flat_map(|i| (0..3).map(move |k| (i, k))).map(|(i, k)| p.jacobian()[(i, k)]).collect();
        let hi: Vec<f32> = (0..3).flat_map(|i| (0..3).map(move |k| (i, k))).map(|(i, k)| p.inverse_hessian()[(i, k)]).collect();
        writeln!(out, "patch_b 96 72 64 {} {} : {} | {} | {}", bits(x), bits(y), dump(p.datas()), dump(&j), dump(&hi)).unwrap();
    }
    std::fs::write(&path, out).expect("cannot write the vector file");
    eprintln!("wrote {}", path);
}
